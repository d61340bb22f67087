#!/bin/bash
# tools/sass_hist.sh <object> <kernel-name-regex> [n]: compact SASS listing (/tmp/k.lst) + opcode histogram of one kernel
OBJ=$1; PAT=$2
cuobjdump -sass $OBJ 2>/dev/null | awk -v pat="$PAT" '/Function : /{p=($0 ~ pat)} p{print}' > /tmp/k.sass
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/k.sass | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*;.*//; s/\s+/ /g' > /tmp/k.lst
wc -l < /tmp/k.lst
awk '{op=$2; if(op ~ /^@/) op=$3; split(op,a,"."); print a[1]}' /tmp/k.lst | sort | uniq -c | sort -rn | head -${3:-30}
