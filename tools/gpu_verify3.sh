#!/bin/bash
TAG=${1:-r01x}; O=gpurun_out; mkdir -p $O
timeout 40 python -m pytest "tests/test_gpu_elementwise.py::test_pack_weights_table_matches_per_tensor_packs" "tests/test_gpu_network.py::test_backward_on_shallow_graphs" "tests/test_gpu_network.py::test_rtod_train_step_against_oracle" "tests/test_gpu_network.py::test_dtod_train_step_against_oracle" "tests/test_gpu_network.py::test_inference_matches_reference_golden" -x -q -m gpu > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
