#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_checkpoint.py -m gpu -q > gpurun_out/pytest_ckpt.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ckpt.log
tail -8 gpurun_out/pytest_ckpt.log
bash tools/ab.sh cfg2 500 default hybird_b200/_ab/lib_LDCS.so hybird_b200/_ab/lib_STCS.so hybird_b200/_ab/lib_LDCSDLB_STCS.so default > gpurun_out/ab_hints.log 2>&1
cat gpurun_out/ab_hints.log
bash tools/ab.sh cfg4 200 default hybird_b200/_ab/lib_LDCSDLB_STCS.so >> gpurun_out/ab_hints.log 2>&1
tail -8 gpurun_out/ab_hints.log
