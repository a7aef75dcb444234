bash scripts/ncu_r02.sh s32
bash scripts/ncu_r02.sh list
