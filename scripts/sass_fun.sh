#!/bin/bash
# SASS of one kernel of the library with decoded scheduling control bits.  usage: sass_fun.sh <substring of mangled name> [regex]
LIB=${LIB:-/root/repo/modulated-deform-conv_b200/lib/libmdconv_b200.so}
NAME=$(cuobjdump -elf $LIB 2>/dev/null | grep -o "\.text\._Z[A-Za-z0-9_]*" | sed 's/^\.text\.//' | grep "$1" | sort -u | head -1)
echo "# $NAME" >&2
cuobjdump -sass -fun "$NAME" $LIB 2>/dev/null | python /root/repo/scripts/sass_ctrl.py "${2:-.}"
