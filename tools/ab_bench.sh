#!/bin/bash
# A/B of two builds of the library over several workloads: tools/ab_bench.sh "<so list>" "<workload list>"
for w in $2; do for so in $1; do
  HUFFB200_SO=$PWD/huff_encoding_b200/$so python bench.py --workload $w --steps 10 --warmup 3 --no-general 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print('$w $so', round(d['value'],1),'GB/s', round(d['ms_per_step'],3),'ms', {n:round(v['ms'],3) for n,v in k.items()})"
done; done
