"""Regenerates the kLowerRuns table of veloci_b200/csrc/format/unicode.hpp.

Every scalar >= U+0080 whose Python `str.lower()` is a single scalar is listed
as (first, last, step, delta) runs; prints the table body to stdout.
"""
pairs = []
for cp in range(0x80, 0x20000):
    c = chr(cp)
    low = c.lower()
    if low != c and len(low) == 1:
        pairs.append((cp, ord(low)))
runs = []
i = 0
while i < len(pairs):
    cp, lo = pairs[i]
    d = lo - cp
    j = i
    while j + 1 < len(pairs) and pairs[j + 1][0] == pairs[j][0] + 1 and pairs[j + 1][1] - pairs[j + 1][0] == d:
        j += 1
    k = i
    while k + 1 < len(pairs) and pairs[k + 1][0] == pairs[k][0] + 2 and pairs[k + 1][1] - pairs[k + 1][0] == d:
        k += 1
    if (k - i) > (j - i):
        runs.append((cp, pairs[k][0], 2, d))
        i = k + 1
    else:
        runs.append((cp, pairs[j][0], 1, d))
        i = j + 1
print(",\n".join("    {0x%X, 0x%X, %d, %d}" % r for r in runs))
