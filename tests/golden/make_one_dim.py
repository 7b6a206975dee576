"""Extracts the 1 x 649 count row of the reference's NaN-guard test
(scan-rs/src/normalization.rs:477-516) into tests/golden/one_dim_649.txt.
Run in the build container only (needs /root/reference); the .txt is committed."""
import re, os
src = open("/root/reference/scan-rs/src/normalization.rs").read()
body = src[src.index("fn test_one_dim()"):]
body = body[body.index("vec!["):body.index("],\n        )")]
nums = [int(x) for x in re.findall(r"\d+", body)]
assert len(nums) == 649, len(nums)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "one_dim_649.txt")
open(out, "w").write(" ".join(map(str, nums)) + "\n")
print("wrote", out)
