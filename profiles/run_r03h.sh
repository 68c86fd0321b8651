for lib in shipped build/variants/libso3d_c3t6.so build/variants/libso3d_c4t4.so; do
  if [ "$lib" = shipped ]; then unset SO3D_LIB_PATH; else export SO3D_LIB_PATH=$lib; fi
  echo "== $lib"; python tests/tools/probe_loop_lanes.py 24 100 | grep '"lanes": "2"'
done
