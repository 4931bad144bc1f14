"""Inline corpora and field configs of the reference's integration tests, ported
verbatim (JSON documents) or translated from TOML to JSON (field configs).

Sources (relative to the reference checkout):
  TEST_ALL      tests/all/tests.rs:9-243       (config :12-39, documents :47-243)
  TEST_SCORE    tests/all/test_scores.rs:6-66
  TEST_PHRASE   tests/all/test_phrase.rs:5-35
  TEST_FACET    tests/all/tests_facet.rs:6-58
  TEST_MINIMAL  tests/all/tests_minimal.rs
"""

BOOST = {"boost": {"boost_type": "f32"}}

TEST_ALL_CONFIG = {
    "*GLOBAL*": {"features": ["All"]},
    "commonness": {"facet": True, **BOOST},
    "ent_seq": {"fulltext": {"tokenize": True}},
    "nofulltext": {"fulltext": {"tokenize": False}},
    "tags[]": {"facet": True},
    "field1[].rank": dict(BOOST),
    "field1[].text": {},
    "kanji[].text": {},
    "meanings.ger[]": {"fulltext": {"tokenize": True}},
    "meanings.eng[]": {"fulltext": {"tokenize": True}},
    "kanji[].commonness": dict(BOOST),
    "kana[].commonness": dict(BOOST),
}

LONG = "Prolog:\nthis is a story of a guy who went out to rule the world, but then died. the end"

# tests/all/tests.rs:41: token values added to the test index after its creation
TEST_ALL_TOKEN_VALUES = ([{"text": "Begeisterung", "value": 20}], {"path": "meanings.ger[]"})

TEST_ALL_DOCS = [
    {"ignore_field": "", "commonness": 123456, "ent_seq": "99999", "tags": ["nice", "cool"]},
    {
        "nofulltext": "my tokens",
        "commonness": 20,
        "tags": ["nice", "cool"],
        "kanji": [{"text": "偉容", "commonness": 0}, {"text": "威容", "commonness": 5}],
        "kana": [{"text": "いよう", "romaji": "Iyou", "commonness": 5}],
        "meanings": {
            "eng": ["karlo", "dignity", "majestic appearance", "will testo"],
            "ger": ["majestätischer Anblick (m)", "majestätisches Aussehen (n)", "Majestät (f)"],
        },
        "ent_seq": "1587680",
    },
    {
        "commonness": 20,
        "tags": ["nice"],
        "kanji": [{"text": "意欲", "commonness": 40}, {"text": "意慾", "commonness": 0}],
        "kana": [{"text": "いよく", "romaji": "Iyoku", "commonness": 40}],
        "meanings": {"eng": ["will", "urge", "having a long torso"], "ger": ["Wollen (n)", "Wille (m)", "Begeisterung (f)", "begeistern"]},
        "ent_seq": "1587690",
    },
    {"meanings": {"eng": ["karl der große"]}},
    {
        "id": 1234566,
        "gender": "male",
        "tags": ["awesome", "cool"],
        "sinlge_value_multi": ["wert"],
        "birthDate": "1960-08-19",
        "address": [{"line": ["nuts strees"]}, {"line": ["asdf"]}],
        "commonness": 500,
        "kanji": [{"text": "意慾", "commonness": 20}],
        "field1": [{"text": "awesome", "rank": 1}],
        "kana": [{"text": "いよく"}],
        "meanings": {"eng": ["test1"], "ger": ["der test", "das ist ein guter Treffer"]},
        "ent_seq": "1587700",
    },
    {
        "id": 123456,
        "tags": ["nice", "cool"],
        "gender": "female",
        "birthDate": "1950-08-19",
        "address": [{"line": ["71955 Ilene Brook"]}],
        "commonness": 551,
        "kanji": [{"text": "何の", "commonness": 526}],
        "field1": [{"text": "awesome"}, {"text": "nixhit"}],
        "kana": [{"text": "どの", "romaji": "Dono", "commonness": 25}],
        "meanings": {"ger": ["welch", "guter nicht Treffer", "alle meine Words", "text", "localität"]},
        "ent_seq": "1920240",
        "mylongtext": LONG,
    },
    {
        "pos": ["adj-i"],
        "commonness": 1,
        "misc": [],
        "tags": ["nice", "cool", LONG],
        "kanji": [{"text": "柔らかい", "commonness": 57}],
        "kana": [{"text": "やわらかい", "romaji": "Yawarakai", "commonness": 30}],
        "meanings": {"ger": ["(1) weich", "stopword"]},
        "ent_seq": "1605630",
    },
    {"meanings": {"ger": ["(1) 2 3 super nice weich"]}, "ent_seq": "9555"},
    {"meanings": {"ger": ["text localität", "alle meine Words"]}, "ent_seq": "1000"},
    {
        "sub_level": [{"text": "Prolog:\nthis is story of a guy who went out to rule the world, but then died. the end"}],
        "commonness": 515151,
        "ent_seq": "25",
        "tags": ["nice", "cool"],
    },
    {"title": "Die Erbin die Sünde", "type": "taschenbuch"},
    {"title": "Die Erbin", "type": "taschenbuch"},
    {"commonness": 30, "title": "COllectif", "meanings": {"ger": ["boostemich"]}},
    {"commonness": 30, "float_value": 5.123, "ent_seq": "26", "tags": ["nice", "coolo"]},
    {"commonness": 20, "ent_seq": "27", "my_bool": True, "tags": ["Eis", "cool"]},
    {"commonness": 20, "ent_seq": "28", "tags": ["nice", "cool"]},
]

TEST_SCORE_CONFIG = {
    "title": {"fulltext": {"tokenize": True}},
    "meanings.ger[].boost": dict(BOOST),
    "meanings.ger[].text": {"fulltext": {"tokenize": True}},
    "commonness": dict(BOOST),
    "order": dict(BOOST),
}
TEST_SCORE_DOCS = [
    {"id": 1, "order": 500, "title": "greg tagebuch 05"},
    {"id": 2, "order": 20, "title": "and some some text 05 this is not relevant let tagebuch greg"},
    {"id": 3, "order": 1000, "title": "greg tagebuch"},
    {"id": 4, "commonness": 41, "meanings": {"ger": [{"text": "Fernsehen-Schauen (n)", "boost": 20}]}},
    {"id": 5, "commonness": 551, "meanings": {"ger": ["welch"]}},
    {"id": 6, "commonness": 2, "meanings": {"ger": ["weich"]}},
]

TEST_PHRASE_CONFIG = {
    "title": {"features": ["Search", "PhraseBoost", "BoostTextLocality"], "fulltext": {"tokenize": True}},
    "tags[]": {"features": ["Search", "PhraseBoost", "BoostTextLocality"], "fulltext": {"tokenize": True}},
}
TEST_PHRASE_DOCS = [
    {"title": "die erbin"},
    {"title": "erbin", "tags": ["die", "erbin"]},
    {"tags": ["greg tagebuch 05"]},
    {"tags": ["greg tagebuch", "05"]},
    {"title": "greg tagebuch", "tags": ["greg tagebuch", "05"]},
]

TEST_FACET_CONFIG = {
    "*GLOBAL*": {"features": ["All"]},
    "tags[]": {"facet": True, "features": ["Facets"]},
    "commonness": {"facet": True},
}
TEST_FACET_DOCS = [
    {
        "commonness": 20,
        "tags": ["nice", "cool"],
        "meanings": {
            "eng": ["karlo", "dignity", "majestic appearance", "will testo"],
            "ger": ["majestätischer Anblick (m)", "majestätisches Aussehen (n)", "Majestät (f)"],
        },
    },
    {"commonness": 20, "tags": ["nice"], "meanings": {"eng": ["will", "urge", "having a long torso"], "ger": ["Wollen (n)", "Wille (m)", "Begeisterung (f)", "begeistern"]}},
    {"commonness": 123456, "tags": ["nice", "cool"]},
    {"meanings": {"eng": ["test1"], "ger": ["der test", "das ist ein guter Treffer"]}},
    {"commonness": 20, "tags": ["Eis", "cool"]},
]

# ---- tests/all/test_query_generator.rs:10-137 (the corpus of the query-generator tests)
TEST_QG_CONFIG = {
    "*GLOBAL*": {"features": ["All"]},
    "commonness": {"facet": True, **BOOST},
    "ent_seq": {"fulltext": {"tokenize": True}},
    "nofulltext": {"fulltext": {"tokenize": False}},
    "tags[]": {"facet": True},
    "field1[].rank": dict(BOOST),
    "field1[].text": {"tokenize": True},
    "kanji[].text": {"tokenize": True},
    "meanings.ger[]": {"stopwords": ["stopword"], "fulltext": {"tokenize": True}},
    "meanings.eng[]": {"fulltext": {"tokenize": True}},
    "kanji[].commonness": dict(BOOST),
    "kana[].commonness": dict(BOOST),
}
TEST_QG_DOCS = [
    {"commonness": 123456, "ent_seq": "99999", "tags": ["nice", "cool"]},
    {
        "ent_seq": "1337",
        "commonness": 20,
        "tags": ["nice", "cool", "ent_seq:99999"],
        "kanji": [{"text": "偉容", "commonness": 0}, {"text": "威容", "commonness": 5}],
        "kana": [{"text": "いよう", "romaji": "Iyou", "commonness": 5}],
        "meanings": {"eng": ["will testo"], "ger": ["majestätischer Anblick (m)", "majestätisches Aussehen (n)", "Majestät (f)"]},
    },
    {
        "ent_seq": "1587690",
        "commonness": 20,
        "tags": ["nice"],
        "kanji": [{"text": "意欲", "commonness": 40}, {"text": "意慾", "commonness": 0}],
        "kana": [{"text": "いよく", "romaji": "Iyoku", "commonness": 40}],
        "meanings": {"eng": ["will", "urge", "having a long torso"], "ger": ["Wollen (n)", "Wille (m)", "Begeisterung (f)", "begeistern"]},
    },
    {"id": 1234566, "tags": ["awesome", "cool"], "commonness": 500, "kanji": [{"text": "意慾", "commonness": 20}], "kana": [{"text": "いよく"}], "ent_seq": "1587700"},
    {"commonness": 515151, "ent_seq": "25", "tags": ["nice", "cool"]},
    {"commonness": 30, "title": "COllectif", "meanings": {"ger": ["boostemich"]}},
    {"commonness": 30, "float_value": 5.123, "ent_seq": "26", "tags": ["nice", "coolo"]},
    {"commonness": 20, "ent_seq": "27", "my_bool": True, "tags": ["Eis", "cool"]},
    {"commonness": 20, "ent_seq": "28", "tags": ["nice", "cool"]},
]

# ---- tests/all/test_why_found.rs:7-64 (the corpus of the why_found tests)
TEST_WHYFOUND_CONFIG = {
    "*GLOBAL*": {"features": ["All"]},
    "richtig": {"fulltext": {"tokenize": True}},
    "not_tokenized": {"fulltext": {"tokenize": False}},
    "not_tokenized_1_n[]": {"fulltext": {"tokenize": False}},
    "custom_tokenized": {"fulltext": {"tokenize": True, "tokenize_on_chars": ["§", "<"]}},
    "url": {"fulltext": {"tokenize": True, "tokenize_on_chars": ["/", ":", "."]}},
}
TEST_WHYFOUND_DOCS = [
    {"url": "https://github.com/PSeitz/veloci", "richtig": "schön super", "viele": ["nette", "leute"]},
    {"not_tokenized": "ID1000", "not_tokenized_1_n": ["ID1000"], "custom_tokenized": "test§_ cool _", "richtig": "hajoe genau"},
    {
        "not_tokenized": "ID2000",
        "not_tokenized_1_n": ["ID2000"],
        "richtig": "shön",
        "custom_tokenized": "<<cool>>",
        "viele": ["treffers", "und so", "super treffers", "ein längerer Text, um zu checken, dass da nicht umsortiert wird"],
    },
    {"buch": "Taschenbuch (kartoniert)", "viele": ["super treffers"]},
]


# ---- tests/all/test_code_search.rs:11-41
TEST_CODE_CONFIG = {
    "*GLOBAL*": {"features": ["All"]},
    "filepath": {"fulltext": {"tokenize": True, "tokenize_on_chars": ["/", "\\"]}},
    "filename": {"fulltext": {"tokenize": True}},
    "line": {"fulltext": {"tokenize": True}},
    "line_number": dict(BOOST),
}
TEST_CODE_DOCS = [{"line_number": 1, "line": "function myfun(param1: Type1)", "filename": "cool.ts", "filepath": "all/the/path"}]
