mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python tools/profile_plan.py unetpp 32 256 400 > gpurun_out/r02v_$tag.txt 2>&1
  echo "$tag: $(grep -E '24<-24 acc=0' gpurun_out/r02v_$tag.txt | head -2 | awk '{print $1}' | tr '\n' ' ') | $(grep -E '48<-48 acc=0' gpurun_out/r02v_$tag.txt | head -2 | awk '{print $1}' | tr '\n' ' ')"
}
run f0 MTBC_FUSE_INBWD=0
run f0_cta1 MTBC_FUSE_INBWD=0 MTBC_HALO_CTAS=1
run f1 MTBC_FUSE_INBWD=1
run f1_noy MTBC_FUSE_INBWD=1 MTBC_HALO_DBG=16
run f1_nomath MTBC_FUSE_INBWD=1 MTBC_HALO_DBG=32
run f1_noboth MTBC_FUSE_INBWD=1 MTBC_HALO_DBG=48
run f1_epi2_noboth MTBC_FUSE_INBWD=1 MTBC_HALO_DBG=48 MTBC_HALO_EPI=2
