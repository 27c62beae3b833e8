"""B200-native NL-Kalman filter / RTS-smoother step (drop-in for the per-frame path of
pariasm/bwd-nlkalman).  The product is libnlkalman_b200.so (CUDA, sm_100a) behind the C
ABI declared in include/; this package is its thin Python binding plus the synthetic
scene generator shared by tests and bench."""
from .api import (FLT1, FLT2, SMO1, Context, NlkError, Params, StripPlan, Tvl1Params, default_params,  # noqa: F401
                  device_count, lib, strip_plan, tvl1_scales, nlkalman_filter_frame, nlkalman_smooth_frame, opp2rgb, rgb2opp, warp_bicubic)
