for o in "5=4" "5=2" "5=1" "5=4" "5=2"; do
MCL_DEBUG=1 MCL_OPTIONS="$o" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/cl_err.txt | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$o', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(r['frac'],4), round(r['kernel_ms_per_step'],3), round(r['fixup_ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
grep "co-resident" gpurun_out/cl_err.txt | sort -u
done
