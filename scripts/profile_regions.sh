#!/bin/bash
# Appends warp-stall reasons and per-region instruction/sample shares of the k_tile launches of
# gpurun_out/<tag>_full.ncu-rep to profiles/<tag>_summary.md. Line ranges follow loops_tile.cuh / sph_math.cuh.
TAG=$1
R="loops_tile.cuh:63-100=mbarrier waits / TMA issue,loops_tile.cuh:146-245=exact sorted-axis path,loops_tile.cuh:248-281=prologue,loops_tile.cuh:282-531=producer,loops_tile.cuh:532-562=consumer setup,loops_tile.cuh:563-656=drain (list merge + exact frames),loops_tile.cuh:657-752=task switch (target loads),loops_tile.cuh:753-809=stage loop + octet cull,loops_tile.cuh:810-878=prefilter test loop,loops_tile.cuh:879-990=flush,sph_math.cuh:60-86=exact r2 / dsubf helpers,sph_math.cuh:87-108=kernel_deval,sph_math.cuh:109-140=sqrt/rcp helpers,sph_math.cuh:141-174=iact_density,sph_math.cuh:175-213=iact_gradient,sph_math.cuh:214-310=iact_force"
{
  echo
  echo "# $TAG: warp-stall reasons (ncu raw page) and where the issued instructions / stall samples fall (SASS samples of the source page joined with nvdisasm -g line info; scripts/ncu_stalls.py, scripts/ncu_lines.py)"
  echo
  echo '```'
  python scripts/ncu_stalls.py gpurun_out/${TAG}_full.ncu-rep | grep -E "##|issue_active|warps_active|stalls"
  echo
  for sel in "k_tile<(int)0, (int)0, (int)4" "k_tile<(int)0, (int)0, (int)2" "k_tile<(int)2"; do
    REGIONS="$R" python scripts/ncu_lines.py gpurun_out/${TAG}_full.ncu-rep "$sel" 0 0 | grep -E "kernel|region" | grep -v "ins   0\.[0-4]"
    echo
  done
  echo '```'
  echo "Note: the source page sums all ncu replay passes, so spin instructions of the mbarrier waits are over-represented there; smsp__inst_executed.sum above is the single-pass count."
} >> profiles/${TAG}_summary.md
