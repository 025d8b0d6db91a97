timeout 600 python -m pytest tests -m gpu -x -q -k "rx_parity or full_size" > gpurun_out/v2_tests.log 2>&1; tail -2 gpurun_out/v2_tests.log
timeout 200 python bench.py --steps 8 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('ms_per_step %.3f front %.3f value %.0f'%(d['ms_per_step'],d['kernels_ms']['rx_front'],d['value']))"
QPSK_B200_LIB=$PWD/tools/bin/libq_prof.so timeout 300 python tools/front2_prof.py 2>&1 | tail -12
