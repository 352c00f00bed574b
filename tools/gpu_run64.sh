#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r6r_pytest.log 2>&1; tail -3 gpurun_out/r6r_pytest.log
python bench.py > gpurun_out/r6r_bench1.json 2> gpurun_out/r6r_bench1.err; echo "rc $?"; tail -2 gpurun_out/r6r_bench1.err
python -c "import __graft_entry__ as g; g.smoke()"
