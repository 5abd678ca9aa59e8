mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python tools/profile_plan.py unetpp 32 256 400 > gpurun_out/r02o_$tag.txt 2>&1
  head -8 gpurun_out/r02o_$tag.txt | sed -n 1,8p
}
run base X=1
run ctas1 MTBC_HALO_CTAS=1
run wcap150 MTBC_HALO_WCAP_KB=150
run widebn128 MTBC_HALO_WIDE_BN=128
run wcap150_widebn128 MTBC_HALO_WCAP_KB=150 MTBC_HALO_WIDE_BN=128
