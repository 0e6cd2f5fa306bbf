#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1
echo "== kernel_roofline exit $?"; grep -E "attention" gpurun_out/kernel_roofline.log
