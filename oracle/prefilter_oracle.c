void orc_prefilter_placeholder(void) {}
