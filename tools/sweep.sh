#!/bin/bash
# usage: sweep.sh "cfg M tpb minblocks maxblocks" ...   (one bench line each)
for spec in "$@"; do
  set -- $spec
  python bench.py --no-cpu-baseline --steps 2 --workload $1 --instances $2 --tpb $3 --minblocks $4 --maxblocks $5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$1 M=$2 tpb=$3 minb=$4 maxb=$5 | blocks', c['grid_blocks'], 'regs', c['registers_per_thread'], 'local', c['local_bytes_per_thread'], 'failed', c['failed_instances'], '| value %.4g e2e %.4g ms %.2f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))"
done
