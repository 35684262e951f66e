#!/bin/bash
# the whole GPU suite without stopping at the first failure
out=gpurun_out; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q > $out/${1:-run}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $out/${1:-run}_pytest.log
