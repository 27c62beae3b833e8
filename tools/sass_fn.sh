#!/bin/bash
# dump the SASS of one kernel of the library (substring match on the mangled name), one instruction per line
# usage: tools/sass_fn.sh <mangled-name-substring> [lib]
LIB=${2:-bwd_nlkalman_b200/libnlkalman_b200.so}
cuobjdump -sass "$LIB" | awk -v pat="$1" '/Function :/{on = index($0, pat) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's/\/\* 0x[0-9a-f]* \*\///'
