tools/bin/packed_peak_bench > gpurun_out/r02_packed_peak.jsonl 2>&1
QPSK_B200_LIB=$PWD/tools/bin/libq_prof.so timeout 300 python tools/front_prof.py > gpurun_out/r02_front_timeline_v1.txt 2>&1
QPSK_B200_FRONT=2 QPSK_B200_LIB=$PWD/tools/bin/libq_prof.so timeout 300 python tools/front2_prof.py > gpurun_out/r02_front2_phases.txt 2>&1
python -c "
from qpsk_b200 import capi
print('probe exact', capi.probe_fp32(0, False), 'fast', capi.probe_fp32(0, True))" > gpurun_out/r02_probe.txt 2>&1
QPSK_B200_FRONT=2 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_front2.log 2>&1; tail -3 gpurun_out/r02_tests_front2.log
cat gpurun_out/r02_probe.txt
