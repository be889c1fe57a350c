import sys
sys.path.insert(0, '/root/repo')
from tests import util
cores, b, q1, q2, _ = util.make_case(60000, 100, seed=77)
o = util.run_oracle(cores, b, q1, q2, bucket_set_bytes=1 << 21)
for w in (8, 5):
    ranks = util.run_sharded_loopback(cores, b, q1, q2, w, bucket_set_bytes=1 << 21)
    print(w, [r[1].stats["rounds"] for r in ranks][:1], [r[1].stats["split"] for r in ranks][:1])
    util.assert_sharded_same(o, ranks)
    print("ok", w)
