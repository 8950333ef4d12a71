LogoData = LogoFormat = LogoOptions = Alphabet = classic = png_print_formatter = formatters = None
