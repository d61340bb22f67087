#!/bin/bash
# tools/sass_spills.sh <object> <mangled-name-substring>: local-memory instructions of one kernel with two lines of context (no GPU needed)
OBJ=$1; PAT=$2
FN=$(cuobjdump -elf "$OBJ" 2>/dev/null | grep -o "_ZN[A-Za-z0-9_]*" | grep "$PAT" | sort -u | head -1)
cuobjdump -sass -fun "$FN" "$OBJ" | grep -E "^\s+/\*[0-9a-f]{4,}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*;.*//' > /tmp/sass_body.txt
grep -n -B2 -A2 "LDL\|STL" /tmp/sass_body.txt
