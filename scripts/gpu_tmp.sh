#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/t_pytest.log 2>&1; tail -14 gpurun_out/t_pytest.log
