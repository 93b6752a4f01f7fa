#!/bin/bash
# config C's shape on one GPU: products under different row-block schedules
run() { python bench.py --no-dense --no-config-d --config-e 0 --steps 5 "$@" 2>>gpurun_out/t16_thin.err | python -c "
import sys, json
l = json.loads(sys.stdin.readline())['spmv']
print('$*', '| fwd %.2f ms (%.3f) trans %.2f ms (%.3f) lsqr %.2f ms/it asm %.1fs' % (l['forward']['ms'], l['forward']['moved_frac'], l['transposed']['ms'], l['transposed']['moved_frac'], l['lsqr']['ms_per_it'], l['assemble_s']), l.get('row_blocks', {}).get('stations_per_block'), l.get('note'))"; }
run
