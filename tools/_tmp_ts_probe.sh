for ko in 16 512 0; do
  echo "== VAR 5 ko $ko"; CLBM_HCZ3D_SWEEP_VAR=5 timeout 120 python tools/hcz3d_sweep_ko.py 16 3 $ko 2>&1 | tail -2
done
