#!/bin/bash
# usage: scratch/sass.sh <object basename> <kernel name substring> -> /tmp/t/<name>.sass
B=/root/repo/homonim_b200/csrc/build
F=$(cuobjdump -sass $B/$1.o | grep -o "_ZN[^ ]*$2[^ ]*" | head -1)
mkdir -p /tmp/t
cuobjdump -sass -fun "$F" $B/$1.o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > /tmp/t/out.sass
wc -l /tmp/t/out.sass
