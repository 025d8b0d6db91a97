QPSK_B200_LIB=$PWD/tools/bin/libq_prof.so timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_front.log 2>&1
grep -c ROLE gpurun_out/prof_front.log
