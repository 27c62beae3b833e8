#!/bin/bash
# where a per-frame CLI invocation spends its time (C1 files), with and without extra hardware queues
OUT=gpurun_out; D=$(mktemp -d)
python - <<PY
import numpy as np, sys
sys.path.insert(0, ".")
from bwd_nlkalman_b200 import synth
w, h, s = 854, 480, 20.0
def pfm(p, a):
    open(p, "wb").write(f"Pf\n{w} {h}\n-1\n".encode() + np.ascontiguousarray(a, np.float32).tobytes())
for t in range(2): pfm("$D/n%d.pfm" % t, synth.noisy_frame(w, h, 1, t, s))
open("$D/bflo.flo", "wb").write(b"PIEH" + np.array([w, h], np.int32).tobytes() + synth.backward_flow(w, h).tobytes())
open("$D/occ.pgm", "wb").write(f"P5\n{w} {h}\n255\n".encode() + synth.occlusion_mask(w, h).astype(np.uint8).tobytes())
PY
B=bwd_nlkalman_b200/bin/nlkalman-flt
$B -i $D/n0.pfm -s 20 --flt11 $D/a1.pfm --flt21 $D/a2.pfm
for rep in 1 2 3; do
  env NLK_CLI_TIMING=1 $B -i $D/n1.pfm -s 20 -o $D/bflo.flo -k $D/occ.pgm --flt10 $D/a1.pfm --flt20 $D/a2.pfm --flt11 $D/b1.pfm --flt21 $D/b2.pfm
done
echo "--- with CUDA_DEVICE_MAX_CONNECTIONS=32"
env CUDA_DEVICE_MAX_CONNECTIONS=32 NLK_CLI_TIMING=1 $B -i $D/n1.pfm -s 20 -o $D/bflo.flo -k $D/occ.pgm --flt10 $D/a1.pfm --flt20 $D/a2.pfm --flt11 $D/b1.pfm --flt21 $D/b2.pfm
echo "--- reference"
time oracle/_ref/nlkalman-flt-ref -i $D/n1.pfm -s 20 -o $D/bflo.flo -k $D/occ.pgm --flt10 $D/a1.pfm --flt20 $D/a2.pfm --flt11 $D/r1.pfm --flt21 $D/r2.pfm
