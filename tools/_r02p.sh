mkdir -p gpurun_out
run() { arch=$1; tag=$2; shift 2
  env "$@" timeout 300 python tools/profile_plan.py $arch 32 256 400 > gpurun_out/r02p_${arch}_$tag.txt 2>&1
  head -2 gpurun_out/r02p_${arch}_$tag.txt
}
run unetpp base X=1
run unetpp old MTBC_HALO_STATS2CTA=1
run nnunet base X=1
run nnunet old MTBC_HALO_STATS2CTA=1
run nnunet wcap150 MTBC_HALO_WCAP_KB=150
run nnunet widebn128 MTBC_HALO_WIDE_BN=128
run nnunet ctas1 MTBC_HALO_CTAS=1
