for lib in ab/librt_r1.so ab/librt_cur.so; do RT_B200_LIB=$lib timeout 100 python scripts/dbg_slow.py car_only 1920 1080; RT_B200_LIB=$lib timeout 100 python scripts/dbg_slow.py car_boxed 1920 1080 traversal=2 ctas_per_sm=8; done > gpurun_out/r2_dbg_slow.log 2>&1
cat gpurun_out/r2_dbg_slow.log
for lib in r1 cur; do RT_B200_LIB=ab/librt_$lib.so timeout 300 ncu --set full --clock-control none -k regex:render_kernel -s 20 -c 1 -o gpurun_out/r2_dbg_$lib -f python scripts/dbg_slow.py car_only 1920 1080 > gpurun_out/r2_dbg_ncu_$lib.log 2>&1; done
ls -la gpurun_out/r2_dbg_*.ncu-rep
