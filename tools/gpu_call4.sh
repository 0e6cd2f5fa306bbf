#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_attn_ab.sh libgraft_x_v3.so
bash tools/gpu_prof_fwd.sh
