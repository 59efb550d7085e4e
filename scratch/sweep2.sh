for f in 8 4 12; do
RBQ_FLUSH_AT=$f python bench.py --steps 8 --warmup 3 --no-cpu-baseline --nprobe 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('flush_at', $f, round(d['value']), round(d['e2e']['value']), d['stage_ms_per_step'], d['scan_split']['ms_head'], d['per_query'])"
done
