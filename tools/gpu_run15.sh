for v in "X=1" "SRB_WG_TILEW=8" "SRB_WG_TILEW=4" "SRB_WG_TILEW=16 SRB_WG_SKIP=1" "SRB_WG_TILEW=16 SRB_WG_SKIP=2"; do
echo "### $v"; env $v python tools/trace_tapgemm.py 2>&1 | grep -A1 "== wgrad" | cut -c1-260
done
