#!/bin/bash
# packets -> PCM throughput (e2e) over the decoder's pipeline chunk length and the step length; run on the GPU box
for T in 32 64; do for C in 8 16 32; do
  timeout 300 python bench.py --frames $T --chunk-frames $C --steps 12 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('frames', $T, 'chunk', $C, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'k_parse_ms', round(d['e2e']['k_parse_ms_per_launch'],3), 'boundary', round(d['e2e_boundary']['value']))"
done; done | tee gpurun_out/chunk_sweep_r2.txt
