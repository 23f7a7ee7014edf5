#!/bin/bash
# static SASS instruction counts of the cooperative building blocks (compile-only, no GPU needed)
set -e
cd "$(dirname "$0")/.."
LAYOUT=${1:-Wide16}
echo "layout $LAYOUT"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPROBE_LAYOUT=$LAYOUT -cubin -o /tmp/coop_probe.cubin tools/coop_sass_probe.cu
for k in probe_gather probe_mulred1 probe_sbox1 probe_sbox3 probe_mds; do
  echo "== $k"
  cuobjdump -sass /tmp/coop_probe.cubin | awk -v k="$k" '/Function :/ {on = index($0, k) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" \
    | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn \
    | awk '{t += $1; printf "%s:%s ", $2, $1} END {printf "\n   total %d\n", t}'
done
