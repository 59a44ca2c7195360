#!/bin/bash
cd "$(dirname "$0")/.."
tools/ab_geo.sh g128l g128s g256s g512s g512l 2>&1 | tee gpurun_out/ab_geo_r02a.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_r02a.txt
