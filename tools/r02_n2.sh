# usage (GPU box): bash tools/r02_n2.sh <tag> -- A/B of the environment toggles in $ROMIS_AB_TOGGLES (full frame, 135-row band), parity subset
R=$1
bash tools/r02_ab.sh $R tests 2>&1 | grep -v "lib_\*"
