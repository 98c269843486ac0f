// Empty stand-in: the reference includes this Boost header but never instantiates an archive
// (SURVEY.md section 8c).  Written for this repo; not a copy of Boost.
#pragma once
