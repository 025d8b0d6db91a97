set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "rx_parity or full_size or api_errors" > gpurun_out/t11_tests.log 2>&1; tail -3 gpurun_out/t11_tests.log
for T in 0 1; do
QPSK_BENCH_TRANSIENT=$T timeout 200 python bench.py --no-cpu-baseline --no-e2e --no-configs > gpurun_out/t11_bench_T$T.json 2>&1
QPSK_BENCH_TRANSIENT=$T ncu --set full --clock-control none -k regex:rx_front -s 3 -c 1 -o gpurun_out/r02_rx_front_v8_T$T -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/ncu_front_T$T.log 2>&1
done
