#!/bin/bash
# per-frame CLI wall time: with / without the GPU warm-up thread, against the resident driver on one frame
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
S=bwd_nlkalman_b200/bin/nlkalman-seq
A="-i $D/n1.pfm -s 20 -o $D/bflo.flo -k $D/occ.pgm --flt10 $D/a1.pfm --flt20 $D/a2.pfm --flt11 $D/b1.pfm --flt21 $D/b2.pfm"
$B -i $D/n0.pfm -s 20 --flt11 $D/a1.pfm --flt21 $D/a2.pfm
T() { local t0=$(date +%s.%N); "$@" > /dev/null; local t1=$(date +%s.%N); echo "wall $(echo "$t1 - $t0" | bc -l 2>/dev/null || python -c "print($t1 - $t0)") s"; }
for rep in 1 2 3; do echo "--- warm-up thread (default)"; T env NLK_CLI_TIMING=1 $B $A; done
for rep in 1 2 3; do echo "--- NLK_CLI_WARMUP=0"; T env NLK_CLI_WARMUP=0 NLK_CLI_TIMING=1 $B $A; done
for rep in 1 2 3; do echo "--- nlkalman-seq, one frame"; T $S -i $D/n%d.pfm -f 0 -l 0 -s 20 --filt1 $D/q1_%d.pfm; done
for rep in 1 2; do echo "--- nlkalman-occ (smallest GPU program)"; T bwd_nlkalman_b200/bin/nlkalman-occ $D/bflo.flo 0.75 $D/o.pgm; done
echo "--- reference"; T oracle/_ref/nlkalman-flt-ref $A
