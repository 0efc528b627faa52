for o in "ew_ctas_per_sm=3" "ew_ctas_per_sm=4" "ew_ctas_per_sm=5" "ew_ctas_per_sm=6"; do
echo "== $o"; AB_OPTS=$o python profiles/scripts/slab_floor.py 8 1 2>&1 | grep slab
done
