for leg in 0 1; do
if [ $leg = 1 ]; then export RBQ_LEGACY_STREAM=1; else unset RBQ_LEGACY_STREAM; fi
RBQ_TRACE=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --nprobe 16 2>gpurun_out/trace_l$leg.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('legacy', $leg, round(d['value']), round(d['e2e']['value']), d['stage_ms_per_step'])"
grep "rbq trace" gpurun_out/trace_l$leg.err | tail -2
done
