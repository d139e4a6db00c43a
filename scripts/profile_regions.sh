#!/bin/bash
# Appends warp-stall reasons and per-region instruction/sample shares of the k_pipe launches of
# gpurun_out/<tag>_full.ncu-rep to profiles/<tag>_summary.md. Line ranges follow loops_pipe.cuh / loops_common.cuh / sph_math.cuh.
TAG=$1
R="loops_pipe.cuh:160-560=producer (task prefetch + fragment layout + TMA issue),loops_pipe.cuh:561-600=consumer setup,loops_pipe.cuh:601-700=drain (list entries -> interactions),loops_pipe.cuh:701-805=task switch (target loads),loops_pipe.cuh:806-855=stage loop + octet cull,loops_pipe.cuh:856-880=target frame coordinates,loops_pipe.cuh:881-935=exact test loop,loops_pipe.cuh:936-1013=flush,loops_common.cuh:130-200=mbarrier waits / TMA issue,loops_common.cuh:201-324=exact sorted-axis path,sph_math.cuh:60-86=r2 / dsubf helpers,sph_math.cuh:87-108=kernel_deval,sph_math.cuh:109-140=sqrt/rcp helpers,sph_math.cuh:141-185=iact_density,sph_math.cuh:186-225=iact_gradient,sph_math.cuh:226-330=iact_force"
{
  echo
  echo "# $TAG: warp-stall reasons (ncu raw page) and where the issued instructions / stall samples fall (SASS samples of the source page joined with nvdisasm -g line info; scripts/ncu_stalls.py, scripts/ncu_lines.py)"
  echo
  echo '```'
  python scripts/ncu_stalls.py gpurun_out/${TAG}_full.ncu-rep | grep -E "##|issue_active|warps_active|stalls"
  echo
  for sel in "k_pipe<(int)0, (int)0, (int)7" "k_pipe<(int)0, (int)0, (int)5" "k_pipe<(int)1" "k_pipe<(int)2"; do
    REGIONS="$R" python scripts/ncu_lines.py gpurun_out/${TAG}_full.ncu-rep "$sel" 0 0 | grep -E "kernel|region" | grep -v "ins   0\.[0-4]"
    echo
  done
  echo '```'
  echo "Note: the source page sums all ncu replay passes, so spin instructions of the mbarrier waits are over-represented there; smsp__inst_executed.sum above is the single-pass count."
} >> profiles/${TAG}_summary.md
