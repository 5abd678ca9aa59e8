mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python tools/profile_plan.py unetpp 32 256 400 > gpurun_out/r02t_$tag.txt 2>&1
  grep -E "dgrad|in_bwd" gpurun_out/r02t_$tag.txt | head -4
}
run f0 MTBC_FUSE_INBWD=0
run f1 MTBC_FUSE_INBWD=1
run f1_epi2 MTBC_FUSE_INBWD=1 MTBC_HALO_EPI=2
run f1_cta2 MTBC_FUSE_INBWD=1 MTBC_HALO_CTAS=2 MTBC_HALO_EPI=2
