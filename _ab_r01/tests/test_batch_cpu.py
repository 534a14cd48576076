"""Replica-mode host logic: case sharding over ranks and the shard merge (no GPU)."""
import importlib
import json

batch = importlib.import_module("01-lbm-2d_b200.batch")


def test_shard_is_a_partition_of_the_sorted_case_list():
    names = [f"case_{i:02d}" for i in (5, 3, 9, 0, 7, 1, 8, 2, 6, 4)]
    parts = [batch.shard(names, r, 3) for r in range(3)]
    assert parts[0] == ["case_00", "case_03", "case_06", "case_09"]
    assert sorted(sum(parts, [])) == sorted(names) and sum(len(p) for p in parts) == len(names)


def test_merge_shards_is_atomic_and_complete(tmp_path):
    for r in range(2):
        (tmp_path / f"sim_results.rank{r}.json").write_text(json.dumps({f"c{r}": {"status": "Success", "final_steps": 10 * (r + 1)}}))
    merged = batch.merge_shards(str(tmp_path), world=3)   # rank 2 produced nothing: tolerated
    assert set(merged) == {"c0", "c1"}
    assert json.loads((tmp_path / "sim_results.json").read_text()) == merged
    assert not (tmp_path / "sim_results.json.tmp").exists()


def _fake_runner(log, fail=(), unstable=()):
    def run(name, cfg, mask, out_dir, device, max_steps, progress):
        log.append(name)
        if name in fail:
            raise RuntimeError("boom")
        open(f"{out_dir}/{name}.npz", "w").close()
        return {"status": "Aborted" if name in unstable else "Success", "reason": "NaN" if name in unstable else None,
                "final_steps": 7}
    return run


def test_resume_plan_follows_the_reference_rules():
    ok, skip = batch.resume_plan(["a", "b", "c", "d"], {"a": "Success", "b": "Failed", "c": "Running"})
    assert ok == 1 and skip == {"a", "b"}          # Running is retried, unknown is run (batch_run.py:78-116)


def test_run_cases_records_status_and_resumes(tmp_path):
    out = str(tmp_path)
    cases = {f"c{i}": ({}, None) for i in range(6)}
    log = []
    res = batch.run_cases(cases, out, runner=_fake_runner(log, fail={"c1"}, unstable={"c4"}))
    assert log == sorted(cases)
    assert res["c0"]["status"] == "Success" and res["c0"]["final_steps"] == 7
    assert res["c1"]["status"] == "Failed" and "boom" in res["c1"]["reason"]
    assert res["c4"]["status"] == "Failed" and "NaN" in res["c4"]["reason"]      # case_executor.py:105-107
    assert not (tmp_path / "c4.npz").exists() and (tmp_path / "c0.npz").exists()  # failed outputs are removed
    merged = batch.merge_shards(out, 1)
    assert {k: v["status"] for k, v in merged.items()} == {"c0": "Success", "c1": "Failed", "c2": "Success",
                                                           "c3": "Success", "c4": "Failed", "c5": "Success"}
    # second session: nothing left to do, Failed is not retried
    log2 = []
    assert batch.run_cases(cases, out, runner=_fake_runner(log2)) == {} and log2 == []
    # a case left Running by a crash is retried; the others keep their records
    merged["c2"] = {"status": "Running"}
    (tmp_path / "sim_results.json").write_text(json.dumps(merged))
    (tmp_path / "sim_results.rank0.json").unlink()
    log3 = []
    res3 = batch.run_cases(cases, out, runner=_fake_runner(log3))
    assert log3 == ["c2"] and res3["c2"]["status"] == "Success"
    assert batch.merge_shards(out, 1)["c1"]["status"] == "Failed"


def test_running_is_written_before_the_case_starts(tmp_path):
    seen = {}

    def runner(name, cfg, mask, out_dir, device, max_steps, progress):
        seen[name] = json.loads((tmp_path / "sim_results.rank0.json").read_text())[name]["status"]
        return {"status": "Success", "final_steps": 1}

    batch.run_cases({"a": ({}, None)}, str(tmp_path), runner=runner)
    assert seen == {"a": "Running"}


def test_max_success_quota_is_split_over_the_ranks(tmp_path):
    cases = {f"c{i}": ({}, None) for i in range(10)}
    logs = [[], []]
    for r in range(2):
        batch.run_cases(cases, str(tmp_path), rank=r, world=2, max_success=5, runner=_fake_runner(logs[r]))
    assert len(logs[0]) == 3 and len(logs[1]) == 2           # 5 successes in total, then stop
    merged = batch.merge_shards(str(tmp_path), 2, remove=True)
    assert sum(v["status"] == "Success" for v in merged.values()) == 5
    # next session with the same quota: already reached, nothing runs (batch_run.py:201-213)
    log = []
    batch.run_cases(cases, str(tmp_path), rank=0, world=2, max_success=5, runner=_fake_runner(log))
    assert log == []


def test_unmerged_shards_of_a_crashed_session_are_carried_over(tmp_path):
    cases = {f"c{i}": ({}, None) for i in range(4)}
    batch.run_cases(cases, str(tmp_path), rank=0, world=2, runner=_fake_runner([]))   # c0, c2 done, never merged
    log = []
    batch.run_cases(cases, str(tmp_path), rank=0, world=1, runner=_fake_runner(log))  # new session consolidates first
    assert log == ["c1", "c3"]
    assert set(batch.merge_shards(str(tmp_path), 1)) == set(cases)


def test_concurrent_cases_overlap_and_keep_the_bookkeeping(tmp_path):
    import threading
    import time

    cases = {f"c{i:02d}": ({}, None) for i in range(12)}
    live, peak, lock = [0], [0], threading.Lock()

    def runner(name, cfg, mask, out_dir, device, max_steps, progress):
        with lock:
            live[0] += 1
            peak[0] = max(peak[0], live[0])
        time.sleep(0.05)
        with lock:
            live[0] -= 1
        if name == "c05":
            raise RuntimeError("boom")
        return {"status": "Success", "final_steps": 3}

    res = batch.run_cases(cases, str(tmp_path), runner=runner, concurrency=4)
    assert peak[0] == 4 and set(res) == set(cases)
    assert res["c05"]["status"] == "Failed" and sum(v["status"] == "Success" for v in res.values()) == 11
    on_disk = json.loads((tmp_path / "sim_results.rank0.json").read_text())
    assert {k: v["status"] for k, v in on_disk.items()} == {k: v["status"] for k, v in res.items()}
    # quota under concurrency: never more successes than asked for, and failures do not eat the quota
    out2 = tmp_path / "q"
    res2 = batch.run_cases(cases, str(out2), runner=runner, concurrency=4, max_success=7)
    assert sum(v["status"] == "Success" for v in res2.values()) == 7
