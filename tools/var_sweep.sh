#!/bin/bash
# `value` of every library variant in build/var/ (built with `make KFLAGS=... OUT=build/var/lib_x.so`); run on the GPU box only:
# the variant is copied over the product library of the box's snapshot.  Writes gpurun_out/var_sweep.txt
ARGS=${ARGS:---steps 10 --warmup 3 --no-e2e --no-cpu-baseline}
cp ffmpeg-heaac_b200/libheaac_b200.so /tmp/lib_orig.so
for f in build/var/*.so; do
  cp $f ffmpeg-heaac_b200/libheaac_b200.so
  timeout 300 python bench.py $ARGS 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value']))" \
    >> gpurun_out/var_sweep.txt
done
cp /tmp/lib_orig.so ffmpeg-heaac_b200/libheaac_b200.so
cat gpurun_out/var_sweep.txt
