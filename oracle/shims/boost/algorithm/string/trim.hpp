// oracle/shims: empty stand-in; the reference only names boost::algorithm::trim inside commented-out code.
#pragma once
