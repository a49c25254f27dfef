"""Stand-in for python-Levenshtein where it is not installed: predicate_alignment.py only calls
Levenshtein.ratio (predicate_alignment.py, name similarity of predicates).

ratio(a, b) = (len(a) + len(b) - d) / (len(a) + len(b)) with d the edit distance in which a
substitution costs 2 (= insertions + deletions only), i.e. d = len(a) + len(b) - 2 * LCS(a, b);
two empty strings give 1.0.  Same definition as python-Levenshtein's ratio()."""


def _lcs(a, b):
    if len(a) < len(b):
        a, b = b, a
    prev = [0] * (len(b) + 1)
    for ca in a:
        cur = [0]
        for j, cb in enumerate(b, 1):
            cur.append(prev[j - 1] + 1 if ca == cb else max(prev[j], cur[j - 1]))
        prev = cur
    return prev[-1]


def ratio(a, b):
    total = len(a) + len(b)
    if total == 0:
        return 1.0
    return 2.0 * _lcs(a, b) / total


def distance(a, b):
    """the classic edit distance (substitution cost 1)"""
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]
