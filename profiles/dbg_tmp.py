import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np
import waldboost_b200 as wb
from waldboost_b200 import synthetic as S
M = wb.Model.load(os.path.join(ROOT, "tests/golden/configA_model.pb"))
frame = S.synthetic_frame(1000, 480, 640)
dt = M.detect(frame)
print(len(dt), M.n_loc, M.n_weak)
