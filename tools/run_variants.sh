for v in "$@"; do
  L=tools/bin/libq_$v.so; [ $v = base ] && L=qpsk_b200/libqpsk_b200.so
  QPSK_B200_LIB=$PWD/$L timeout 200 python bench.py --steps 8 --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v ms_per_step %.3f front %.3f value %.0f'%(d['ms_per_step'],d['kernels_ms']['rx_front'],d['value']))"
done
