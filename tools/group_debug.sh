#!/bin/bash
# debug helper: the overlay's BEAM DD example on one and on two devices of one process, monitor rows side by side
B=ddpca-admm_b200/host/_bin/beam_dd_b200
for musc in 2 3; do
  for devs in 0 0,1; do
    d=$(mktemp -d); (cd $d && DDPCA_DEVICES=$devs $OLDPWD/$B --glob 1 --doma 8,1,1 --musc $musc | cut -c1-200)
    echo "musc=$musc devices=$devs:"; awk '{print NR-1, $(NF-1), $NF}' $d/Beam/resuMoni.txt
  done
done
