cp variants/lib_w80.so object_slam_b200/libobslam_b200.so
timeout 120 python -m pytest tests/test_gpu_extractor.py -m gpu -x -q 2>&1 | tail -1
cp variants/lib_base.so object_slam_b200/libobslam_b200.so
timeout 400 bash tools/ab.sh 2 base w80 2>&1 | grep -E "^(base|w80)"
