/* Hand-written stand-in for the bison-generated header of the reference's
 * nutmeg expression grammar (bison is not installed here). Test infrastructure. */
#ifndef NGB200_PP_TOKENS_H
#define NGB200_PP_TOKENS_H
enum { TOK_NUM = 258, TOK_STR = 259, TOK_LE = 260, TOK_GE = 261, TOK_NE = 262 };
typedef union YYSTYPE {
    double num;
    const char *str;
    struct pnode *pnode;
} YYSTYPE;
#endif
