#!/bin/bash
# Tuning aid: time the cascade launch for several prebuilt variants of the library (variants/libwbg_<name>.so, built
# with different -D flags) on the same frames.  usage: profiles/variant_sweep.sh FRAMES name1 name2 ...
frames=$1; shift
echo '[{"name": "default"}]' > /tmp/one_variant.json
cp waldboost_b200/libwbg.so /tmp/libwbg_orig.so
for v in "$@"; do
  cp variants/libwbg_$v.so waldboost_b200/libwbg.so
  echo "== $v"
  python profiles/cascade_sweep.py $frames /tmp/one_variant.json 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print({k: d[k] for k in ('ok', 'kernel_ms_per_frame', 'hits', 'eval_cost', 'exec_per_live')}, [round(x, 3) for x in d['exec_by_nk']])
print('checksum', d['hits'], d['eval_cost'])"
done
cp /tmp/libwbg_orig.so waldboost_b200/libwbg.so
