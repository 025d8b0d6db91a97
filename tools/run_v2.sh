timeout 600 python -m pytest tests -m gpu -x -q -k "rx_parity or full_size or api_errors" > gpurun_out/v2_tests.log 2>&1; tail -5 gpurun_out/v2_tests.log
for v1 in 0 1; do
QPSK_B200_FRONT_V1=$v1 timeout 200 python bench.py --steps 8 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('v1=$v1 ms_per_step %.3f front %.3f value %.0f'%(d['ms_per_step'],d['kernels_ms']['rx_front'],d['value']))"
done
