#!/bin/bash
# wide (nfft 1024) TX: parity tests of the one-pass kernel, then the bench line of both paths
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wide_tx_resident or wideband_1024" > gpurun_out/wtx_tests.log 2>&1
tail -5 gpurun_out/wtx_tests.log
OFDM_TX_PATH=twopass timeout 300 python bench.py --workload tx --nfft 1024 --syms 128 --steps 20 > gpurun_out/wtx_twopass.json 2> gpurun_out/wtx_twopass.err
cat gpurun_out/wtx_twopass.json | cut -c1-900
timeout 300 python bench.py --workload tx --nfft 1024 --syms 128 --steps 20 > gpurun_out/wtx_resident.json 2> gpurun_out/wtx_resident.err
cat gpurun_out/wtx_resident.json | cut -c1-900
tail -3 gpurun_out/wtx_resident.err
