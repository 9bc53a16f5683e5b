#!/bin/bash
# round 2, step 10 (the last GPU minutes): interpreter with liveness-based physical slots -- plan + tpch tests, plan-only microbench
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_plan.py tests/test_gpu_tpch.py -q -m gpu -x 2>&1 | tail -3
MOB_PROFILE_ONLY=plan timeout 40 python tools/profile_ops.py 2> gpurun_out/r02_plan_v10.err | tee gpurun_out/r02_plan_v10.json | cut -c1-600
