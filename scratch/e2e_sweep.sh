#!/bin/bash
# e2e A/B on one box: chunk count x head placement
for late in 0 1; do for ch in 0 4 8 16; do
  if [ $late = 1 ]; then export RBQ_HEAD_LATE=1; else unset RBQ_HEAD_LATE; fi
  if [ $ch = 0 ]; then unset RBQ_FEED_CHUNKS; else export RBQ_FEED_CHUNKS=$ch; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --nprobe 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('late', $late, 'chunks', $ch, 'device', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
