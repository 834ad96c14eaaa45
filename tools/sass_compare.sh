#!/bin/bash
# tools/sass_compare.sh libA.so libB.so [kernel-name-regex]  -- compares the SASS of every kernel whose (mangled) name matches the
# regex (default: k_pcg5) between two builds, instruction by instruction (encodings stripped).  Used to show that cuba_pcg5.cuh's
# kernels in the final library are the measured ones (profiles/r02_pcg5_big_regression.log).  Needs no GPU.
A=$1; B=$2; RE=${3:-k_pcg5}
norm() { grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##'; }
rc=0
for k in $(cuobjdump -sass "$A" 2>/dev/null | grep "Function :" | awk '{print $3}' | grep -E "$RE" | sort -u); do
	cuobjdump -sass -fun "$k" "$A" 2>/dev/null | norm > /tmp/sass_a.$$
	cuobjdump -sass -fun "$k" "$B" 2>/dev/null | norm > /tmp/sass_b.$$
	na=$(wc -l < /tmp/sass_a.$$); nb=$(wc -l < /tmp/sass_b.$$)
	if [ "$nb" -eq 0 ]; then echo "$k: $na instructions, absent in $B"; continue; fi
	if cmp -s /tmp/sass_a.$$ /tmp/sass_b.$$; then echo "$k: $na instructions, IDENTICAL"; else echo "$k: $na vs $nb instructions, DIFFERENT"; rc=1; fi
done
rm -f /tmp/sass_a.$$ /tmp/sass_b.$$
exit $rc
