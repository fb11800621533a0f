// flare/base/random.h is included by flare/base/experimental/bloom_filter.h but none of
// the members the oracle instantiates use it.
#pragma once
