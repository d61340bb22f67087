#!/bin/bash
# usage: tools/sass_stats.sh <object> <mangled-name-substring>  -- opcode histogram and spill traffic of one kernel (no GPU needed)
OBJ=$1; PAT=$2
FN=$(cuobjdump -elf "$OBJ" 2>/dev/null | grep -o "_ZN[A-Za-z0-9_]*" | grep "$PAT" | sort -u | head -1)
[ -z "$FN" ] && { echo "no function matching $PAT"; exit 1; }
cuobjdump -sass -fun "$FN" "$OBJ" | grep -E "^\s+/\*[0-9a-f]{4,}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' > /tmp/sass_body.txt
echo "$FN: $(wc -l < /tmp/sass_body.txt) instructions"
sed -E 's/^(@!?U?P[0-9]+\s+)?//' /tmp/sass_body.txt | awk '{print $1}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-30} | awk '{printf "%s %s, ", $2, $1} END {print ""}'
