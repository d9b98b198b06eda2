run() { echo "== $*"; env "$@" python tools/stage_ablation.py $SL 1440 base 2>&1 | tail -1; }
SL=12 run ABL_GROUP=4 ABL_CORR_CTAS=296
SL=12 run ABL_GROUP=4 ABL_CORR_CTAS=222
SL=12 run ABL_GROUP=4 ABL_CORR_CTAS=120
SL=16 run ABL_GROUP=8 ABL_CORR_CTAS=148
SL=12 run ABL_GROUP=3 ABL_CORR_CTAS=148
