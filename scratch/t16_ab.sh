#!/bin/bash
run() { python bench.py --no-dense --no-config-d --config-e 0 --steps 5 "$@" 2>>gpurun_out/t16_ab.err | python -c "
import sys, json
l = json.loads(sys.stdin.readline())['spmv']
print('$*', '| fwd %.3f ms (%.3f) trans %.3f ms (%.3f) lsqr %.3f ms/it asm %.1fs' % (l['forward']['ms'], l['forward']['moved_frac'], l['transposed']['ms'], l['transposed']['moved_frac'], l['lsqr']['ms_per_it'], l['assemble_s']), l.get('row_blocks', {}).get('stations_per_block'), l.get('note'))"; }
run
den() { python bench.py --no-compressed --no-config-d --config-e 0 --no-cpu-baseline --steps 10 "$@" 2>>gpurun_out/t16_ab.err | python -c "
import sys, json
l = json.loads(sys.stdin.readline())
print('$*', '| value %.2f it/s sweep %.3f ms frac %.4f clocks %s' % (l['value'], l['roofline']['launch_ms'], l['roofline']['frac'], l['clocks']['sm_mhz']))"; }
den
den --dense-vec4 2
den --dense-vec4 2 --dense-f2f-rows 1
den --dense-vec4 2 --dense-f2f-rows 0
