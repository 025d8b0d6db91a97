QPSK_B200_LIB=$PWD/tools/bin/libq_prof.so timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_front.log 2>&1
grep -c PROF gpurun_out/prof_front.log
for d in 0 10000 20000 30000 0 20000; do
  QPSK_B200_DEPHASE=$d timeout 200 python bench.py --steps 8 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('dephase $d ms_per_step %.3f front %.3f value %.0f'%(d['ms_per_step'],d['kernels_ms']['rx_front'],d['value']))"
done
