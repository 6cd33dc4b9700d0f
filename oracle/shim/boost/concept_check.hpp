#pragma once
#define BOOST_CONCEPT_ASSERT(x) static_assert(true, "")
