"""Protobuf messages of the reference's model file (reference waldboost/model.proto:1-23), built at import
time with the protobuf runtime (there is no protoc in the image and the reference's generated model_pb2.py is
git-ignored).  Field numbers and types are the wire contract: files written by either implementation load in
the other.  Note ChannelOpts.func is field 5; there is no field 4."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _build():
    fd = descriptor_pb2.FileDescriptorProto(name="waldboost_b200/model.proto", syntax="proto3")
    rep, opt = _F.LABEL_REPEATED, _F.LABEL_OPTIONAL
    spec = {
        "Model": [("shape", 1, _F.TYPE_INT32, rep, ""), ("channel_opts", 2, _F.TYPE_MESSAGE, opt, ".ChannelOpts"),
                  ("classifier", 3, _F.TYPE_MESSAGE, rep, ".DTree"), ("theta", 4, _F.TYPE_FLOAT, rep, "")],
        "ChannelOpts": [("shrink", 1, _F.TYPE_INT32, opt, ""), ("n_per_oct", 2, _F.TYPE_INT32, opt, ""),
                        ("smooth", 3, _F.TYPE_INT32, opt, ""), ("func", 5, _F.TYPE_STRING, opt, "")],
        "DTree": [("feature", 1, _F.TYPE_INT32, rep, ""), ("threshold", 2, _F.TYPE_FLOAT, rep, ""),
                  ("left", 3, _F.TYPE_INT32, rep, ""), ("right", 4, _F.TYPE_INT32, rep, ""),
                  ("prediction", 5, _F.TYPE_FLOAT, rep, "")],
    }
    for mname, fields in spec.items():
        m = fd.message_type.add(name=mname)
        for name, num, typ, label, tname in fields:
            f = m.field.add(name=name, number=num, type=typ, label=label)
            if tname:
                f.type_name = tname
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName(n)) for n in spec}


_classes = _build()
Model = _classes["Model"]
ChannelOpts = _classes["ChannelOpts"]
DTree = _classes["DTree"]
