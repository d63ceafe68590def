#!/bin/bash
# host_cli leg with different thread settings
mkdir -p gpurun_out
for io in 4 8 12; do for dt in 8 16; do
TB_IO_THREADS=$io TB_DECODE_THREADS=$dt timeout 600 python bench.py --reads 200000 --cov-records 0 --steps 1 --warmup 1 --no-e2e --cpu-sample 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())['host_cli']; print('io $io dec $dt', round(d['value']/1e6,2), 'M/s wall', round(d['wall_s'],2), 'dec', d['decode_merge_s'], 'pack', d['pack_s'], 'dev', d['device_s'], 'tagwrite', d['tag_write_s'])"
done; done
