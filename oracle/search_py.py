"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product): a second, independent restatement of the hot path in
plain Python + numpy float32 -- request JSON -> get_term_ids_in_field -> resolve_token_to_anchor -> union / intersect ->
add_boost -> top_n -- over the oracle's own decoder of the index files (oracle/index_files.py).  It shares no code with the
product and none with oracle/veloci_oracle.cpp (which reads index directories through the product's reader and parses
requests with the product's DOM parser): tests/test_python_search_oracle.py holds the two oracles against each other.

Covers SURVEY 8 rows a1, a4-a7, a9, a11, a12, a17 (search / or / and trees of search parts with levenshtein_distance,
starts_with, ignore_case, boost; request boosts on the anchor level with every boost function, expression and
skip_when_score; top / skip), and a10 / a13 / a14 / a16: `filter` trees, phrase boosts, text locality, facets -- the
shape of BASELINE config 3 -- plus a8 (boosts on a 1:n level), a15 (boost_term), the per-part top / skip bound and token
values of a4, and suggest / suggest_multi (f.2).  Not covered (raises): regex parts.

Each function cites the reference file:line it follows.
"""
import numpy as np

import index_files as oif
import read_document as ord_

F = np.float32


def _lower_one(ch):  # the matching automaton folds scalar by scalar (no expansion, no context)
    low = ch.lower()
    return low if len(low) == 1 else ch


def _prefix_distances(a, b, transposition):
    """Levenshtein distance (adjacent transposition at cost one when asked) of every prefix of `a` to `b`: [d(a[:i], b)]."""
    n, m = len(a), len(b)
    pp, p = None, list(range(m + 1))
    out = [p[m]]
    for i in range(1, n + 1):
        c = [i] + [0] * m
        for j in range(1, m + 1):
            v = min(p[j] + 1, c[j - 1] + 1, p[j - 1] + (a[i - 1] != b[j - 1]))
            if transposition and i >= 2 and j >= 2 and a[i - 1] == b[j - 2] and a[i - 2] == b[j - 1]:
                v = min(v, pp[j - 2] + 1)
            c[j] = v
        pp, p = p, c
        out.append(p[m])
    return out


def _edit_distance(a, b, transposition):
    return _prefix_distances(a, b, transposition)[-1]


def _distance_u8(s1, s2):  # search_field.rs:705-732: plain DP in u8 cells, 255 when a string has 255 bytes or more
    if len(s1.encode("utf-8")) >= 255 or len(s2.encode("utf-8")) >= 255:
        return 255
    column = list(range(len(s1) + 1))
    for x, cx in enumerate(s2):
        column[0] = (x + 1) & 0xFF
        lastdiag = x & 0xFF
        for y, cy in enumerate(s1):
            if cy != cx:
                lastdiag = (lastdiag + 1) & 0xFF
            olddiag = column[y + 1]
            column[y + 1] = min((column[y + 1] + 1) & 0xFF, (column[y] + 1) & 0xFF, lastdiag)
            lastdiag = olddiag
    return column[len(s1)]


def _default_score(distance, prefix_matches):  # search_field.rs:27-33
    if prefix_matches:
        return F(2.0) / (np.log2(F(distance) + F(1.0)) + F(0.2))
    return F(2.0) / (F(distance) + F(0.2))


class Unsupported(Exception):
    pass


class PySearch:
    def __init__(self, directory):
        self.ix = oif.IndexDirectory(directory)
        self.meta = {ix["path"]: ix for _field, ix in self.ix.indices()}
        self._dicts, self._files = {}, {}
        self.reader = ord_.Reader(directory)  # key-value stores (oracle/read_document.py, over the same decoder)

    def _file(self, name):
        if name not in self._files:
            self._files[name] = self.ix.read(name)
        return self._files[name]

    def _dictionary(self, path):  # [(text, term id)] in key order: the FST stream's order
        if path not in self._dicts:
            if not any(f + ".textindex" == path and c.get("has_fst") for f, c in self.ix.meta["columns"].items()):
                raise KeyError("field does not exist %s (fst not found)" % path)
            self._dicts[path] = [(k.decode("utf-8"), v) for k, v in oif.Fst(self._file(path + ".fst")).items()]
        return self._dicts[path]

    # ---- search_field.rs:277-398
    def field_search(self, part):
        if part.get("is_regex"):
            raise Unsupported("is_regex")
        path = part["path"] if part["path"].endswith(".textindex") else part["path"] + ".textindex"
        term = part["terms"][0]
        lower_term = term.lower()
        d = part.get("levenshtein_distance")
        if d is not None:
            d = min(d, len(lower_term) - 1)  # :286 (the empty term is not exercised here)
        d_score = d or 0
        d_match = min(d_score, 4)                              # :87-88
        ignore_case = part.get("ignore_case")
        transposition = bool(ignore_case) if ignore_case is not None else False  # :87 (sic)
        case_insensitive = ignore_case if ignore_case is not None else True
        starts_with = bool(part.get("starts_with"))
        check_prefix = starts_with or d_score != 0             # :302
        limit = part.get("top") is not None                    # :292-294
        top_n = (part.get("top") if limit else 10) + (part.get("skip") or 0)
        worst = F(-3.40282347e+38)
        order = lambda h: (-float(h[1]), -h[0])                # search.rs:123-130: score desc, id desc
        query = [(_lower_one(c) if case_insensitive else c) for c in term]
        hits = []
        for text, term_id in self._dictionary(path):           # the FST stream: ascending key = ascending term id
            cand = [(_lower_one(c) if case_insensitive else c) for c in text]
            if starts_with:  # the automaton's prefix closure: some prefix of the key is within distance
                ok = min(_prefix_distances(cand, query, transposition)) <= d_match
            else:
                ok = abs(len(cand) - len(query)) <= d_match and _edit_distance(cand, query, transposition) <= d_match
            if not ok:
                continue
            line_lower = text.lower()
            prefix_matches = check_prefix and line_lower.startswith(lower_term)
            k = _edit_distance(list(line_lower), list(lower_term), True)   # distance_dfa :691-702
            dist = k if k <= d_score else _distance_u8(line_lower, lower_term)
            score = _default_score(dist, prefix_matches)
            if limit:                                          # :322-331, sort.rs:25-34
                if score < worst:
                    continue
                if hits and len(hits) == top_n + 200:
                    hits.sort(key=order)
                    del hits[top_n:]
                    if hits:
                        worst = hits[-1][1]
            hits.append((term_id, score))
        if part.get("boost") is not None:                      # :359-364
            hits = [(i, s * F(part["boost"])) for i, s in hits]
        if limit:                                              # :366-369 (the reference's sort is unstable: ties at the cut are open)
            hits.sort(key=lambda h: -float(h[1]))
            del hits[top_n:]
        if part.get("token_value"):                            # :391-395: add_boost over the term hits
            tv = dict(part["token_value"])
            tv["path"] = tv["path"] + ".textindex.token_values"
            scored = dict(hits)
            self.add_boost(tv, scored)
            hits = [(i, scored[i]) for i, _ in hits]
        return path, hits

    # ---- search_field.rs:400-464
    def resolve(self, path, hits):
        name = path + ".to_anchor_id_score"
        meta = self.meta[name]
        start_pos, data = self._file(name + ".indirect"), self._file(name + ".data")
        best = {}
        for term_id, score in hits:
            for anchor, raw in oif.anchor_scores(start_pos, data, term_id, wide=meta.get("data_type") == "U64"):
                final = score * (F(np.float16(F(raw))) / F(100.0))  # :426: the stored score passes through f16
                if anchor not in best or final > best[anchor]:
                    best[anchor] = final
        return best

    # ---- set_op.rs:87-220
    @staticmethod
    def union(results):
        if len(results) == 1:
            return results[0]
        slots = sorted({term for term, _ in results})
        out = {}
        for anchor in sorted(set().union(*[hits.keys() for _, hits in results])):
            mx = [F(0.0)] * len(slots)
            for term, hits in results:
                if anchor in hits:
                    i = slots.index(term)
                    mx[i] = max(mx[i], hits[anchor])
            n = F(sum(1 for m in mx if m >= F(0.00001)))
            total = F(0.0)
            for m in mx:
                total = total + m
            out[anchor] = total * n * n
        return (results[0][0], out)

    # ---- set_op.rs:368-446
    @staticmethod
    def intersect(results):
        if len(results) == 1:
            return results[0]
        lists = list(results)
        shortest = min(range(len(lists)), key=lambda i: (len(lists[i][1]), i))   # the first shortest
        short = lists[shortest][1]
        lists[shortest] = lists[-1]                                              # swap_remove
        lists.pop()
        out = {}
        for anchor in sorted(short):
            if all(anchor in hits for _, hits in lists):
                score = F(0.0)
                for _, hits in lists:
                    score = score + hits[anchor]
                out[anchor] = score + short[anchor]
        return (lists[0][0], out)

    def _tree(self, node, boosts=()):
        if "search" in node:
            part = node["search"]
            path, term_hits = self.field_search(part)
            hits = self.resolve(path, term_hits)
            pos = part["path"].rfind("[]")
            if pos >= 0:  # plan_creator_search_part (execution_plan.rs:422-509): a boost on the part's own 1:n level
                level = part["path"][:pos]
                on_level = [b for b in boosts if b["path"].rfind("[]") >= 0 and b["path"][:b["path"].rfind("[]")] == level]
                if len(on_level) > 1:
                    raise ValueError("more than one boost on the same 1:n level")
                if on_level:
                    self._apply_anchor_boost(hits, on_level[0], self._boost_to_anchor(part, path, term_hits, on_level[0]))
            return (part["terms"][0], hits)
        kind = "or" if "or" in node else "and"
        inputs = []
        for q in node[kind]["queries"]:
            own = ((q.get("search") or q.get("or") or q.get("and")).get("options") or {}).get("boost") or []  # merge_vec :263-270
            inputs.append(self._tree(q, tuple(boosts) + tuple(own)))
        return self.union(inputs) if kind == "or" else self.intersect(inputs)

    # ---- BoostToAnchor (plan_steps.rs:174-196): term hits -> text ids -> value ids -> (anchor, boost value) in value id order
    def _boost_to_anchor(self, part, path, term_hits, boost):
        field = path[:-len(".textindex")]
        tokenized = (((self.ix.meta["columns"].get(field) or {}).get("textindex_metadata") or {}).get("options") or {}).get("tokenize")
        ids = [i for i, _ in term_hits]
        if tokenized:  # resolve_token_hits_to_text_id_ids_only (search_field.rs:640-689): a token without an entry is itself a text id
            text_ids = []
            for i in ids:
                vals = self.reader.get_values(path + ".tokens_to_text_id", i)
                text_ids.extend(vals if vals is not None else [i])
            ids = sorted(set(text_ids))
        value_ids = sorted({v for i in ids for v in (self.reader.get_values(path + ".value_id_to_parent", i) or [])})  # join_to_parent_ids (search.rs:281-315)
        out = []
        for value_id in value_ids:  # get_boost_ids_and_resolve_to_anchor (boost.rs:432-468)
            v = self._boost_value(boost["path"], value_id)
            if v is None:
                continue
            anchor = self.reader.get_value(boost["path"] + ".value_id_to_anchor", value_id)
            if anchor is not None:
                out.append((anchor, v))
        return out

    # ---- ApplyAnchorBoost: apply_boost_values_anchor (boost.rs:255-281), the merge walk as it is written
    def _apply_anchor_boost(self, hits, boost, boost_ids):
        if not boost_ids:
            return
        it = iter(boost_ids)
        cur = next(it)
        for anchor in sorted(hits):
            if cur[0] < anchor:
                for b in it:
                    if b[0] > anchor:
                        cur = b
                        break
                    if b[0] == anchor:
                        cur = b
                        hits[anchor] = self._apply_boost(boost, hits[anchor], b[1])
            elif cur[0] == anchor:
                hits[anchor] = self._apply_boost(boost, hits[anchor], cur[1])

    def _ids_of_part(self, part):  # hits_ids of a part resolved to anchors, one entry per (matched text id, anchor) (search_field.rs:466-498)
        path, hits = self.field_search(part)
        field = path[:-len(".textindex")]
        if self.ix.meta["columns"].get(field, {}).get("is_anchor_identity_column"):
            return [term_id for term_id, _ in hits]
        anchors = []
        for term_id, _ in hits:
            anchors.extend(self.reader.get_values(path + ".text_id_to_anchor", term_id) or [])
        return anchors

    def _ids_tree(self, node):
        if "search" in node:
            return set(self._ids_of_part(node["search"]))
        kind = "or" if "or" in node else "and"
        sets = [self._ids_tree(q) for q in node[kind]["queries"]]
        return set().union(*sets) if kind == "or" else set.intersection(*sets)

    # ---- phrase boosts (execution_plan.rs:202-262, plan_steps.rs:235-293, search_field.rs:263-275, boost.rs:380-402)
    def _phrase_boosts(self, phrase_boosts, hits):
        groups = {}
        for pb in phrase_boosts:
            p1, p2 = pb["search1"], pb["search2"]
            assert p1["path"] == p2["path"]
            path, h1 = self.field_search(p1)
            _, h2 = self.field_search(p2)
            name = path + ".phrase_pair_to_anchor"
            store = oif.phrase_pair_records(self._file(name + ".indirect"), self._file(name + ".data"))
            anchors = groups.setdefault((p1["terms"][0], p2["terms"][0]), set())  # the same two terms in several fields: one group
            for t1, _ in h1:
                for t2, _ in h2:
                    anchors.update(store.get((t1, t2), []))
        for anchors in groups.values():  # every group multiplies the hits it contains by 5.0 (plan_steps.rs:271)
            for a in anchors:
                if a in hits:
                    hits[a] = hits[a] * F(5.0)

    # ---- text locality (boost.rs:11-87, search.rs:113-120,180-184)
    def _term_ids_in_field(self, node, out):
        if "search" in node:
            path, hits = self.field_search(node["search"])
            if hits:  # search_field.rs:379-383
                out.setdefault(path, {})[node["search"]["terms"][0]] = [i for i, _ in hits]
            return
        for q in node["or" if "or" in node else "and"]["queries"]:
            self._term_ids_in_field(q, out)

    def _text_locality(self, search_req, hits):
        per_field = {}
        self._term_ids_in_field(search_req, per_field)
        best = {}
        for path, term_to_ids in per_field.items():
            if len(term_to_ids) <= 1:
                continue
            counts = {}
            for ids in term_to_ids.values():
                for token in ids:
                    for text_id in self.reader.get_values(path + ".tokens_to_text_id", token) or []:
                        counts[text_id] = counts.get(text_id, 0) + 1
            identity = self.ix.meta["columns"].get(path[:-len(".textindex")], {}).get("is_anchor_identity_column")
            for text_id, c in counts.items():
                if c <= 1:
                    continue
                boost = F(2.0) * F(c) * F(c)
                anchors = [text_id] if identity else (self.reader.get_values(path + ".text_id_to_anchor", text_id) or [])
                for a in anchors:
                    if a not in best or boost < best[a]:  # the reversed comparator keeps the minimum (boost.rs:23-28)
                        best[a] = boost
        for a, boost in best.items():
            if a in hits:
                hits[a] = hits[a] * boost

    # ---- facet.rs:31-83
    def facet(self, field, top, anchors):
        steps, cur = [], ""
        for piece in field.split("."):
            cur = cur + "." + piece if cur else piece
            if piece.endswith("[]"):
                steps.append(cur)
        steps.append(field + ".textindex")  # util.rs:173-188 get_steps_to_anchor
        counts = {}
        if len(steps) == 1 or self.reader.has_index(steps[-1] + ".anchor_to_text_id"):
            path = steps[0] + ".parent_to_value_id" if len(steps) == 1 else steps[-1] + ".anchor_to_text_id"
            for a in anchors:
                for v in self.reader.get_values(path, a) or []:
                    counts[v] = counts.get(v, 0) + 1
        else:
            ids = list(anchors)
            for step in steps:  # join_anchor_to_leaf
                ids = [v for i in ids for v in (self.reader.get_values(step + ".parent_to_value_id", i) or [])]
            for v in ids:
                counts[v] = counts.get(v, 0) + 1
        groups = sorted(counts.items(), key=lambda g: -g[1])  # ties: unspecified in the reference (unstable sort)
        if top is not None:
            groups = groups[:top]
        return [(self.reader.text_for_id(steps[-1], v), n, v) for v, n in groups]

    # ---- boost.rs:470-504, 283-377; expression.rs:25-100
    def _boost_value(self, path, anchor):
        name = path + ".boost_valid_to_value"
        meta = self.meta.get(name)
        if meta is None:
            raise KeyError("Did not found path in indices " + name)
        if meta["index_cardinality"] == "MultiValue":
            vals = oif.indirect_get_values(self._file(name + ".indirect"), self._file(name + ".data"), anchor)
            bits = vals[0] if vals else None
        else:
            bits = oif.packed_get_value(self._file(name), oif.packed_width(meta["metadata"]["max_value_id"]), anchor)
        return None if bits is None else np.array([bits], dtype=np.uint32).view(np.float32)[0]

    @staticmethod
    def _expression(text, v):
        ops, current = [], ""

        def flush(s):
            try:
                ops.append(("f", F(float(s))))
            except ValueError:
                pass

        for c in text:
            if c == " ":
                if current:
                    flush(current)
                current = ""
            else:
                current += c
            if current in ("+", "-", "/", "*"):
                ops.append((current, None))
                current = ""
            elif current == "$SCORE":
                ops.append(("$", None))
                current = ""
        if current:
            flush(current)
        left = v if ops[0][0] == "$" else ops[0][1]
        right = v if ops[2][0] == "$" else ops[2][1]
        with np.errstate(divide="ignore", invalid="ignore"):
            op = ops[1][0]
            return left / right if op == "/" else left * right if op == "*" else left + right if op == "+" else left - right

    def _apply_boost(self, boost, score, v):  # apply_boost (boost.rs:283-377)
        param = F(boost.get("param") or 0.0)
        fun = boost.get("boost_fun")
        with np.errstate(divide="ignore", invalid="ignore"):
            if fun == "Log10":
                score = score * np.log10(v + param)
            elif fun == "Log2":
                score = score * np.log2(v + param)
            elif fun == "Multiply":
                score = score * (v + param)
            elif fun == "Add":
                score = score + (v + param)
            elif fun == "Replace":
                score = v + param
            if boost.get("expression"):
                score = score + self._expression(boost["expression"], v)
        return F(score)

    def add_boost(self, boost, hits):  # boost.rs:470-504
        skip = [F(x) for x in boost.get("skip_when_score") or []]
        for anchor in list(hits):
            score = hits[anchor]
            if any(abs(x - score) < F(0.00001) for x in skip):
                continue
            v = self._boost_value(boost["path"], anchor)
            if v is not None:
                hits[anchor] = self._apply_boost(boost, score, v)

    # ---- search_field.rs:160-228: suggest_multi / suggest
    def suggest_multi(self, request):
        items = []
        for part in request["suggest"]:
            path, hits = self.field_search(part)
            texts = {term_id: text for text, term_id in self._dictionary(path)}
            items.extend((texts[term_id].lower(), score, term_id) for term_id, score in hits)  # return_term_lowercase
        items.sort(key=lambda it: it[0], reverse=True)  # (stable where the reference's sort is not: equal texts keep their order)
        merged = []
        for text, score, term_id in items:  # dedup_by: equal texts merge into the first, which takes the larger score
            if merged and merged[-1][0] == text:
                if score > merged[-1][1]:
                    merged[-1] = (text, score, merged[-1][2])
            else:
                merged.append((text, score, term_id))
        merged.sort(key=lambda it: -float(it[1]))
        skip, top = request.get("skip") or 0, request.get("top")
        merged = merged[skip:]
        return [(t, float(s), i) for t, s, i in (merged if top is None else merged[:top])]

    def suggest(self, part):  # :219-228: the part's own top / skip also bound the merged list
        return self.suggest_multi({"suggest": [part], "top": part.get("top"), "skip": part.get("skip")})

    # ---- search.rs:143-228
    def search(self, request):
        for key in ("suggest", "select"):
            if request.get(key):
                raise Unsupported(key)
        top = request.get("top", 10)
        top = 10 if top is None else top
        skip = request.get("skip") or 0
        _, hits = self._tree(request["search_req"], tuple(request.get("boost") or []))
        hits = dict(hits)
        if request.get("filter"):  # intersect_score_hits_with_ids (set_op.rs:311-326)
            allowed = self._ids_tree(request["filter"])
            hits = {a: s for a, s in hits.items() if a in allowed}
        for boost in request.get("boost") or []:
            if "[]" not in boost["path"]:  # execution_plan.rs:175-189: boosts on the anchor level follow the tree
                self.add_boost(boost, hits)
        if request.get("phrase_boosts"):
            self._phrase_boosts(request["phrase_boosts"], hits)
        for part in request.get("boost_term") or []:  # search.rs:176, boost.rs:89-195: hits the part also finds, times its boost (2.0)
            factor = F(part["boost"] if part.get("boost") is not None else 2.0)
            for a in self._ids_of_part(part):  # an anchor reached from two matched texts is boosted twice (boost.rs:222-231)
                if a in hits:
                    hits[a] = hits[a] * factor
        if request.get("text_locality"):
            self._text_locality(request["search_req"], hits)
        ordered = sorted(hits.items(), key=lambda h: (-float(h[1]), -h[0]))  # sort.rs:5-22 / search.rs:123-130: score desc, id desc
        out = {"num_hits": len(hits), "data": [(a, float(s)) for a, s in ordered[skip:skip + top]]}
        if request.get("facets"):  # search.rs:188-206: over the sorted hit ids
            out["facets"] = {f["field"]: self.facet(f["field"], f.get("top", 10), sorted(hits)) for f in request["facets"]}
        return out
