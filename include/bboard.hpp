#include "pom_bboard.hpp"   /* drop-in name: code written against the reference includes "bboard.hpp" */
