#!/usr/bin/env bash
# Developer helper: SASS instruction count per kernel of an object file / library (16 B each; the bounce loop
# has to fit the 32 KB instruction cache).  usage: tools/sass_sizes.sh file.o
cuobjdump -sass "$1" | awk '/Function :/ {if (name) print n, name; name=$3; n=0} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/{n++} END{print n, name}' | sed -E 's/_ZN[0-9A-Za-z_]*(render_pool_kernel|render_kernel)/\1/'
