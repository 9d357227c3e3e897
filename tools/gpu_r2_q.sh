#!/bin/bash
timeout 120 python tools/bench_qrcp.py 4096 2048 f32 2>&1 | tail -1 | cut -c1-300
RLB200_QR_NOCOOP=1 timeout 120 python tools/bench_qrcp.py 4096 2048 f32 2>&1 | tail -1 | cut -c1-300
timeout 120 python tools/bench_qrcp.py 4096 2048 f64 2>&1 | tail -1 | cut -c1-300
RLB200_QR_NOCOOP=1 timeout 120 python tools/bench_qrcp.py 4096 2048 f64 2>&1 | tail -1 | cut -c1-300
timeout 120 python tools/bench_qrcp.py 256 65536 f64 2>&1 | tail -1 | cut -c1-300
RLB200_QR_NOCOOP=1 timeout 120 python tools/bench_qrcp.py 256 65536 f64 2>&1 | tail -1 | cut -c1-300
timeout 120 python tools/bench_qrcp.py 20000 256 f64 0 2>&1 | tail -1 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_cqrrpt.py tests/test_gpu_bqrrp.py tests/test_gpu_drivers.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
