/* pom_agents.hpp — kept for code written against round 1: the policies now live in include/agents.hpp (the drop-in
 * for the reference's header, same names and members, plus seed constructors). */
#ifndef POM_AGENTS_HPP_
#define POM_AGENTS_HPP_
#include "agents.hpp"
#endif
