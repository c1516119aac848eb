#!/bin/bash
# usage: tools/dump_sass.sh <object-or-so> <kernel-name-substring> <out>   -- one "addr OPCODE operands" line per instruction
cuobjdump -sass "$1" 2>/dev/null | awk -v k="$2" '/Function :/{on = index($0, k) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//' > "$3"
wc -l "$3"
