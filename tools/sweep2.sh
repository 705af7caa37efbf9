#!/bin/bash
# usage: sweep2.sh "cfg M lanes tpb" ...
for spec in "$@"; do
  set -- $spec
  python bench.py --no-cpu-baseline --steps 3 --workload $1 --instances $2 --lanes $3 --tpb $4 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$1 M=$2 lanes=$3 tpb=$4 | blocks', c['grid_blocks'], 'regs', c['registers_per_thread'], '| value %.4g ms %.3f' % (d['value'], d['ms_per_step']))"
done
