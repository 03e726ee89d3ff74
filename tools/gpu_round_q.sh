#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_misc.py -m gpu -q -p no:cacheprovider -k val_epoch > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -25 gpurun_out/${TAG}_tests.log
