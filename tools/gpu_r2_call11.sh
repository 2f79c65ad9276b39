#!/bin/bash
mkdir -p gpurun_out
echo "== K1 trace"; timeout 120 python tools/diag_basis_trace.py
