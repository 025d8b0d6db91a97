#!/bin/bash
# Opcode histogram of kernels in libqpsk_b200.so whose mangled name matches a pattern.
#   tools/sass_hist.sh <pattern> [top-N]
# Prints the ptxas resource line and the static SASS opcode counts (predicates stripped).
cd "$(dirname "$0")/../qpsk_b200" || exit 1
pat="$1"; top="${2:-18}"
grep -A3 "Compiling entry function '[^']*${pat}" csrc/ptxas.log | grep -E "Compiling|spill|Used" | paste - - - \
  | sed -E 's/ptxas info    ://g; s/Function properties.*bytes stack frame,//; s/Compiling entry function//' | cut -c1-220
cuobjdump -sass libqpsk_b200.so | awk -v pat="$pat" '/Function : /{p=($0 ~ pat)} p' > /tmp/sass_hist.$$
echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_hist.$$)"
grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_hist.$$ | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/\..*//; s/;//' \
  | sort | uniq -c | sort -rn | head -"$top" | awk '{printf "%s %s  ", $1, $2} END {print ""}'
mv /tmp/sass_hist.$$ /tmp/sass_last.sass
