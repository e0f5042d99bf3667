# experiments: wgrad band shapes (SRB_WG_FORCE="TH,wsplit") with the rows+co-stacked flavour forced (flag 524288)
for f in "0,0" "2,2" "3,2" "2,3" "3,3" "4,3" "3,4" "4,4"; do
  echo "force TH,ws=$f"
  SRB_WG_FORCE=$f WGRAD=524288, timeout 100 python tools/time_rs.py 2>&1 | grep -E "vdsr body|espcn L3|rror"
done
