# ncu --set full of the DP warp kernel on config 4: plain layout (HFA_PAIR=0) vs pair layout (default / forced)
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
for p in ${PAIRS:-0 1}; do
  HFA_PAIR=$p python bench.py --workload c4 --no-cpu --no-extra --steps 2 --warmup 3 > $O/plain_p$p.log 2>&1 && \
  HFA_PAIR=$p ncu --set full --clock-control none --import-source on -k regex:hfa_dp_warp_any -s 2 -c 1 -o $O/prof_c4_p$p -f \
      python bench.py --workload c4 --no-cpu --no-extra --steps 2 --warmup 3 > $O/ncu_p$p.log 2>&1
  tail -1 $O/ncu_p$p.log | cut -c1-160
done
