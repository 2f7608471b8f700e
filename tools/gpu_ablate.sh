for lib in "" t0 t4 t8 tn; do
  if [ -n "$lib" ]; then export FUVS_DEV_LIB=flood_uav_video_segmentation_b200/lib/libfuvs_$lib.so; fi
  echo "lib=$lib"
  for m in dense dense_smooth; do for s in 1 2; do timeout 200 python tools/exp_streams.py $m $s 30 2>&1 | grep streams; done; done
done
