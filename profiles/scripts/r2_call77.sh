#!/bin/bash
# call 77: sanity run of the relinked library (SASS-identical rebuild): grouped-engine and driver tests
timeout 200 python -m pytest tests/test_grp_gpu.py tests/test_drivers_gpu.py -m gpu -q -x 2>&1 | tail -2
