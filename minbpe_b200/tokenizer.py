GPT2_SPLIT_PATTERN = GPT4_SPLIT_PATTERN = None
BasicTokenizer = RegexTokenizer = Tokenizer = get_stats = merge = None
