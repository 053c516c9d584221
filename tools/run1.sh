for nl in 2 4; do
export LPFNT_DEBUG_EVAL_NL=$nl
echo "NL=$nl"
python tools/quick_eval.py 4 20 1e7
python tools/quick_eval.py 3 30 1e7
python tools/quick_eval.py 6 12 1e6
python tools/quick_eval.py 4 20 1e7 chebyshev
python tools/quick_eval.py 2 14 1e7
done
unset LPFNT_DEBUG_EVAL_NL
python tools/overlap.py
LPFNT_DEBUG_PUSH_CTAS=2 python tools/overlap.py
python -m pytest tests -m gpu -x -q -k "eval" 2>&1 | tail -3
