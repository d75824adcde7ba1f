#!/bin/bash
# static SASS instruction count per kernel of the built library.  usage: tools/sass_count.sh [lib]
LIB=${1:-nyxus_b200/libnyxus_b200.so}
cuobjdump -sass $LIB | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {cnt[name]++} END {for (n in cnt) print cnt[n], n}' | c++filt | sort -k2 | cut -c1-120
