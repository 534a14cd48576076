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
