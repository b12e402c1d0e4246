for cfgs in "8 2 1" "8 1 1" "16 2 1" "16 1 1" "16 2 2" "16 1 2" "8 2 2" "4 2 1" "12 2 1"; do
  set -- $cfgs
  WHICH=fwd BEVPOOL_FWD_WARPS=$1 BEVPOOL_FWD_CPW=$2 BEVPOOL_FWD_MINB=$3 python profiles/time_kernels.py 2>&1 | tail -1
done
