python -m pytest tests -m gpu -x -q -k "eval" 2>&1 | tail -3
for tr in 0 1 2; do
export LPFNT_DEBUG_EVAL_TREE=$tr
echo "TREE=$tr"
python tools/quick_eval.py 4 20 1e7
python tools/quick_eval.py 3 30 1e7
[ $tr != 2 ] && python tools/quick_eval.py 6 12 1e6
[ $tr != 2 ] && python tools/quick_eval.py 2 14 1e7
[ $tr != 2 ] && python tools/quick_eval.py 1 20 1e7
[ $tr != 2 ] && python tools/quick_eval.py 3 6 1e7
done
