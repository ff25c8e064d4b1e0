// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for <htslib/hts_log.h>
// (main.cpp:38,253,423 only switch logging off).
#pragma once
enum htsLogLevel { HTS_LOG_OFF = 0, HTS_LOG_ERROR, HTS_LOG_WARNING = 3, HTS_LOG_INFO, HTS_LOG_DEBUG, HTS_LOG_TRACE };
static inline void hts_set_log_level(enum htsLogLevel) {}
