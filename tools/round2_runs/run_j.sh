mkdir -p gpurun_out
export B200_GRAPHS=0
timeout 70 ncu --set full --import-source on --clock-control none -k regex:"k_vp8_token_count|k_vp8_token_write|k_vp8_mbmask" -c 3 -f -o gpurun_out/r2r_vp8tok python tools/profile_legs.py webp > /dev/null 2>&1
ls -la gpurun_out | grep r2r
