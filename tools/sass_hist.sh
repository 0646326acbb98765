#!/bin/bash
# usage: tools/sass_hist.sh <kernel-name-substring>   -> opcode histogram of the kernel in libmfvgp.so
cuobjdump -sass meanfieldvargp_b200/lib/libmfvgp.so | awk -v k="$1" '/Function : /{f=index($0,k)>0} f' > /tmp/k.sass
grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed -E 's/\..*//; s/;//' | sort | uniq -c | sort -rn | awk '{t+=$1; print} END {print t, "TOTAL"}'
