#!/bin/bash
# rebuild the CUDA library and print registers / spills / SASS size per kernel
cd "$(dirname "$0")/.." || exit 1
python -m meshoptimizer_b200.build 2>&1 | grep -E "error|Used|spill|so$"
cuobjdump -sass meshoptimizer_b200/lib/libmeshopt_b200.so > /tmp/sass.txt
awk '/Function : /{name=$3} /^        \/\*[0-9a-f]+\*\/ /{n[name]++} END{for(k in n) print k, n[k], n[k]*16/1024 " KB"}' /tmp/sass.txt
