from densephrases_b200.runtime import get_query2vec, load_phrase_index, load_qa_pairs  # noqa: F401
