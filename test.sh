#!/usr/bin/env bash
# Same entry as the reference's test.sh: decode every image under --testDir (default ./data/val) with the latest snapshot.
set -e
cd "$(dirname "$0")"
exec python -m lstm_ctc_ocr_b200.lstm.test_net --network=LSTM_test --cfg=lstm_ctc_ocr_b200/lstm/lstm.yml --restore=1 "$@"
