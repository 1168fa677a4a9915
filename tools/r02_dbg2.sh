#!/bin/bash
echo "== baseline"; python tools/r02_dbg.py 2>&1 | head -3
echo "== EAGER"; CUDA_MODULE_LOADING=EAGER python tools/r02_dbg.py 2>&1 | head -3
echo "== MAXCONN 32"; CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/r02_dbg.py 2>&1 | head -3
echo "== both"; CUDA_MODULE_LOADING=EAGER CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/r02_dbg.py 2>&1 | head -3
echo "== MAXCONN 1"; CUDA_DEVICE_MAX_CONNECTIONS=1 python tools/r02_dbg.py 2>&1 | head -3
