#!/bin/bash
# value / e2e with and without the programmatic-dependent-launch chain, over batch sizes (writes gpurun_out/pdl_sweep.txt)
for S in ${STREAMS:-1024 2048 4096 8192}; do
  for P in 0 1; do
    HEAAC_B200_PDL=$P timeout 300 python bench.py --streams $S --steps 20 2>&1 | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('streams', $S, 'pdl', $P, 'value', round(d['value']), 'e2e', round(d['e2e']['value']))" \
      >> gpurun_out/pdl_sweep.txt
  done
done
