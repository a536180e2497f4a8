#!/bin/bash
# usage: tools/sass_hist.sh <kernel-name-substring>   -- opcode histogram of one kernel of libmsda_sm100.so
so=${2:-ocpg_b200/lib/libmsda_sm100.so}
cuobjdump -sass "$so" | awk -v pat="$1" '
/Function :/ {on = index($0, pat) > 0}
on && /^ +\/\*[0-9a-f]{4}\*\// {print}' > /tmp/sass_one.txt
echo "instructions: $(wc -l < /tmp/sass_one.txt)"
sed -E 's@/\*[0-9a-f]{4}\*/@@; s@/\* 0x[0-9a-f]+ \*/@@' /tmp/sass_one.txt | awk '{op=$1; if (op ~ /^@/) op=$2; sub(/\..*/, "", op); print op}' | sort | uniq -c | sort -rn | head -${3:-30}
