cp variants/lib_p2.so object_slam_b200/libobslam_b200.so
timeout 300 python -m pytest tests/test_gpu_extractor.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
cp variants/lib_base.so object_slam_b200/libobslam_b200.so
bash tools/ab.sh 2 base p2
