for cfg in "2 2 2" "3 2 2" "1 1 1" "3 1 1"; do
  set -- $cfg
  echo "CFG CAP_FWD=$1 CAP_RED=$2 CAP_APPLY=$3"
  HB_BN_CAP_FWD=$1 HB_BN_CAP_RED=$2 HB_BN_CAP_APPLY=$3 timeout 120 python tools/dev_bn_time.py
done
