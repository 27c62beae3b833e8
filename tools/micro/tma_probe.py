import sys, numpy as np
sys.path.insert(0, '.')
import bwd_nlkalman_b200 as nlk
from bwd_nlkalman_b200 import synth
w, h, ch, sigma = 128, 96, 3, 20.0
f1 = nlk.default_params(sigma, nlk.FLT1)
n0 = nlk.rgb2opp(synth.noisy_frame(w, h, ch, 0, sigma))
a = nlk.nlkalman_filter_frame(n0, None, None, sigma, f1)
print("ok", float(a.mean()))
