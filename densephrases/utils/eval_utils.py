from densephrases_b200.runtime import (drqa_exact_match_score, drqa_metric_max_over_ground_truths, drqa_normalize,  # noqa: F401
                                       drqa_regex_match_score, exact_match_score, f1_score, normalize_answer)
