#!/bin/bash
T=r3r
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/${T}_pytest.log
tail -2 gpurun_out/${T}_pytest.log
timeout 300 python tools/k4_update.py 3 2>&1 | tail -1
