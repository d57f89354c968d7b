#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 20 python -m pytest "tests/test_gpu_elementwise.py::test_pack_weights_table_matches_per_tensor_packs" -x -q -m gpu > $O/r01y_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r01y_pytest.log
tail -3 $O/r01y_pytest.log
