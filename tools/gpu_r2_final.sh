#!/bin/bash
# End-of-round confirmation on one box: every GPU test, smoke, default bench (+ breakdown), reference arm, train bench, train-step launch lists.
bash tools/gpu_confirm.sh
bash tools/gpu_train_lists.sh
