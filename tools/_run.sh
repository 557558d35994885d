for v in drm; do
cp variants/lib_$v.so object_slam_b200/libobslam_b200.so
timeout 300 python -m pytest tests/test_gpu_extractor.py -m gpu -x -q 2>&1 | tail -1
done
cp variants/lib_base.so object_slam_b200/libobslam_b200.so
bash tools/ab.sh 2 base drm
