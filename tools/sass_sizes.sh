#!/bin/bash
# SASS instruction count of every kernel / device function in the given object files
for f in "$@"; do
  cuobjdump -sass $f | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@!]/{cnt[name]++} END{for(n in cnt) printf "%8d  %s\n", cnt[n], substr(n,1,100)}' | sort -rn
done
