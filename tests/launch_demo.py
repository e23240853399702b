"""Trainer stand-in for launcher tests (reference: tests/unittests/launch_demo.py:18-20): logs the
environment contract it was started with, optionally runs for a while, exits with
$PADDLE_DEMO_EXIT_CODE."""
import json
import os
import sys
import time

keys = ["PADDLE_JOB_ID", "PADDLE_POD_ID", "PADDLE_TRAINER_ID", "PADDLE_TRAINER_RANK_IN_POD",
        "PADDLE_TRAINERS_NUM", "PADDLE_TRAINER_ENDPOINTS", "PADDLE_CURRENT_ENDPOINT", "FLAGS_selected_gpus",
        "EDL_POD_LEADER_ID", "EDL_POD_IDS", "EDL_STAGE", "RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"]
rec = {k: os.environ.get(k) for k in keys}
rec["pid"] = os.getpid()
rec["t"] = time.time()
out = os.environ.get("DEMO_RECORD_DIR")
if out:
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "start_%s_%d.json" % (rec["PADDLE_POD_ID"], rec["pid"])), "w") as f:
        json.dump(rec, f)
print("demo trainer", json.dumps(rec), flush=True)
run_s = float(os.environ.get("DEMO_RUN_SECONDS", "0"))
done_flag = os.environ.get("DEMO_DONE_FLAG")
t0 = time.time()
while time.time() - t0 < run_s:
    if done_flag and os.path.exists(done_flag):
        break
    time.sleep(0.1)
sys.exit(int(os.environ.get("PADDLE_DEMO_EXIT_CODE", "0")))
