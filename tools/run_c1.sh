for k in 0 115 0 115 100; do QPSK_B200_LOOP_EXCL_KB=$k timeout 120 python tools/config1_time.py 2>&1 | tail -1; done
sh tools/run_variants.sh poll200 base poll8000 unroll2 poll200 base unroll2
