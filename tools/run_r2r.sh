set -x
for rep in 1 2; do for ne in 1 0; do
  S2S_NO_EARLY_LOADS=$ne python tools/profile_model.py --batch 16 --steps 400 > gpurun_out/r2r_profile_b16_noearly${ne}_rep$rep.log 2>&1; echo "noearly=$ne $(head -1 gpurun_out/r2r_profile_b16_noearly${ne}_rep$rep.log | cut -c100-180)"
done; done
S2S_NO_EARLY_LOADS=1 S2S_NO_BN_FOLD=1 python tools/profile_model.py --batch 16 --steps 400 | head -1 | cut -c100-180
