#!/bin/bash
# nlkalman-seq.sh -- frame-by-frame NL-Kalman filtering and smoothing of a sequence,
# driving the B200 binaries nlkalman-flt / nlkalman-smo.
#
# Same positional interface, same intermediate files and same per-frame commands as
# the reference pipeline (reference scripts/nlkalman-seq.sh:4-12, :31-32, :53-54,
# :117-119), so it can be swapped in:
#
#   nlkalman-seq.sh SEQ FFR LFR SIG OUT [STP] [FPM] [SPM] [OPM]
#     SEQ  noisy frames, printf pattern (e.g. in/%03d.tif)
#     FFR  first frame        LFR  last frame        SIG  noise standard deviation
#     OUT  output folder      STP  frame step (default 1)
#     FPM  extra options for nlkalman-flt (quoted)
#     SPM  extra options for nlkalman-smo (quoted); "no" = skip the smoothing
#     OPM  optical flow parameters "FSCALE1 DW1 TH1 FSCALE2 DW2 TH2"
#
# Outputs in OUT: flt1-%03d.tif flt2-%03d.tif smo1-%03d.tif, plus the flows and
# occlusion masks bflo1-%03d.flo bocc1-%03d.png fflo-%03d.flo focc-%03d.png (kept and
# reused on a re-run, so a sequence can be resumed at any frame).
#
# The optical flow estimator (tvl1flow) and the occlusion mask (nlkalman-occ, standing in for the
# reference's plambda expression) are the package's GPU programs, installed next to this script;
# other builds can be given through the TVL1FLOW / PLAMBDA / NLKALMAN_OCC environment variables.

set -u
SEQ=$1; FFR=$2; LFR=$3; SIG=$4; OUT=$5
STP=${6:-1}
FPM=${7:-""}
SPM=${8:-""}
OPM=${9:-"1 0.25 0.75 1 0.25 0.75"}

HERE=$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd -P)
tool() { # tool NAME ENVVALUE
	if [ -n "$2" ]; then echo "$2"; elif [ -x "$HERE/$1" ]; then echo "$HERE/$1"; else command -v "$1" || echo "$HERE/$1"; fi
}
FLT=$(tool nlkalman-flt "${NLKALMAN_FLT:-}")
SMO=$(tool nlkalman-smo "${NLKALMAN_SMO:-}")
TVL1=$(tool tvl1flow "${TVL1FLOW:-}")
PLAMBDA=$(tool plambda "${PLAMBDA:-}")
OCC=$(tool nlkalman-occ "${NLKALMAN_OCC:-}")

mkdir -p "$OUT"
for i in $(seq "$FFR" "$STP" "$LFR"); do
	f=$(printf "$SEQ" "$i")
	if [ ! -f "$f" ]; then echo "ERROR: $f not found"; exit 1; fi
done

read -ra OF <<< "$OPM"
flt1() { printf "$OUT/flt1-%03d.tif" "$1"; }
flt2() { printf "$OUT/flt2-%03d.tif" "$1"; }
smo1() { printf "$OUT/smo1-%03d.tif" "$1"; }

# flow FROM TO OUTFILE NPROC DW FSCALE   (tvl1flow: nproc tau lambda theta nscales fscale)
flow() { [ -f "$3" ] || "$TVL1" "$1" "$2" "$3" "$4" 0 "$5" 0 0 "$6"; }
# occlusion FLOWFILE TH OUTFILE: |divergence of the flow| > TH, as 0 / 255
# (nlkalman-occ evaluates the reference's plambda expression on the GPU; plambda itself is used
# when it is the only one of the two that is installed)
occlusion() {
	[ -f "$3" ] && return
	if [ -x "$OCC" ]; then "$OCC" "$1" "$2" "$3"
	else "$PLAMBDA" "$1" "x(0,0)[0] x(-1,0)[0] - x(0,0)[1] x(0,-1)[1] - + fabs $2 > 255 *" -o "$3"; fi
}

# ---- filtering, forward in time -------------------------------------------------------
# first frame: both filterings, spatial only
"$FLT" -i "$(printf "$SEQ" "$FFR")" -s "$SIG" $FPM --flt11 "$(flt1 "$FFR")" --flt21 "$(flt2 "$FFR")"

for i in $(seq $((FFR + STP)) "$STP" "$LFR"); do
	p=$((i - STP))
	noisy=$(printf "$SEQ" "$i")
	bflo=$(printf "$OUT/bflo1-%03d.flo" "$i")
	bocc=$(printf "$OUT/bocc1-%03d.png" "$i")
	flow "$noisy" "$(flt2 "$p")" "$bflo" 8 "${OF[1]}" "${OF[0]}"
	occlusion "$bflo" "${OF[2]}" "$bocc"
	# first filtering only (the last --f2_p wins), guided by the previous first filtering
	"$FLT" -i "$noisy" -s "$SIG" $FPM --f2_p 0 -o "$bflo" -k "$bocc" \
		--flt10 "$(flt1 "$p")" --flt11 "$(flt1 "$i")"
	# second filtering only, guided by the previous second filtering
	"$FLT" -i "$noisy" -s "$SIG" $FPM --f1_p 0 -o "$bflo" -k "$bocc" \
		--flt11 "$(flt1 "$i")" --flt20 "$(flt2 "$p")" --flt21 "$(flt2 "$i")"
done

[ "$SPM" == "no" ] && exit 0

# ---- smoothing, backward in time ------------------------------------------------------
cp "$(flt2 "$LFR")" "$(smo1 "$LFR")"
for i in $(seq $((LFR - STP)) -"$STP" "$FFR"); do
	n=$((i + STP))
	fflo=$(printf "$OUT/fflo-%03d.flo" "$i")
	focc=$(printf "$OUT/focc-%03d.png" "$i")
	flow "$(flt2 "$i")" "$(smo1 "$n")" "$fflo" 2 "${OF[4]}" "${OF[3]}"
	occlusion "$fflo" "${OF[5]}" "$focc"
	"$SMO" --flt1 "$(flt2 "$i")" --smo0 "$(smo1 "$n")" -s "$SIG" $SPM -o "$fflo" -k "$focc" \
		--smo1 "$(smo1 "$i")"
done
exit 0
