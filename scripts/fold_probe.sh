for c in cfg2 cfg3; do for ph in 1 2 0; do ESVO_DBG_FOLD_PHASE=$ph python scripts/fold_probe.py $c 2>&1 | tail -1; done; ESVO_FOLD_SORT=0 python scripts/fold_probe.py $c 2>&1 | tail -1; done
