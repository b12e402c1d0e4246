for cfgs in "128 3" "128 4" "64 3" "256 3" "256 4" "512 3" "0 3"; do
  set -- $cfgs
  echo "chunk=$1 minb=$2: $(BEVPOOL_FWD_CHUNK=$1 BEVPOOL_FWD_MINB=$2 python profiles/step_breakdown.py 2>&1 | tail -1 | cut -c40-320)"
done
echo "tile: $(BEVPOOL_FWD_KERNEL=tile python profiles/step_breakdown.py 2>&1 | tail -1 | cut -c40-320)"
echo "occ chunk128: $(python profiles/step_breakdown.py occ_200x200x16_b64 4 2>&1 | tail -1 | cut -c40-320)"
echo "hires chunk128: $(python profiles/step_breakdown.py bevdepth_hires_b16 4 2>&1 | tail -1 | cut -c40-320)"
