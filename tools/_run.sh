cp variants/lib_tma.so object_slam_b200/libobslam_b200.so
timeout 120 python -m pytest tests/test_gpu_extractor.py -m gpu -x -q 2>&1 | grep -E "passed|failed|ObsError" | head -5
cp variants/lib_base.so object_slam_b200/libobslam_b200.so
timeout 300 bash tools/ab.sh 2 base tma 2>&1 | grep -E "^(base|tma)"
