#!/usr/bin/env bash
# Same entry as the reference's train.sh.  One GPU:   ./train.sh
# One process per GPU of this node (NCCL):             GPUS=8 ./train.sh
set -e
cd "$(dirname "$0")"
ARGS="--network=LSTM_train --cfg=lstm_ctc_ocr_b200/lstm/lstm.yml --restore=0 $*"
if [ "${GPUS:-1}" -gt 1 ]; then
  exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 --master-port "${PORT:-29511}" \
       -m lstm_ctc_ocr_b200.lstm.train_net $ARGS
fi
exec python -m lstm_ctc_ocr_b200.lstm.train_net $ARGS
