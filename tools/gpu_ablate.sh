for lib in "" u; do
  if [ -n "$lib" ]; then export FUVS_DEV_LIB=flood_uav_video_segmentation_b200/lib/libfuvs_$lib.so; fi
  echo "lib=$lib"
  for m in block block_lowres; do for s in 1 2; do timeout 200 python tools/exp_streams.py $m $s 30 2>&1 | grep streams; done; done
done
