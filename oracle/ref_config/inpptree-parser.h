/* Hand-written stand-in for the bison-generated header of the reference's
 * B-source parse-tree grammar (bison is not installed here). Test infrastructure. */
#ifndef NGB200_PT_TOKENS_H
#define NGB200_PT_TOKENS_H
enum { TOK_NUM = 258, TOK_STR = 259, TOK_pnode = 260, TOK_LE = 261, TOK_LT = 262,
       TOK_GE = 263, TOK_GT = 264, TOK_EQ = 265, TOK_NE = 266, TOK_OR = 267, TOK_AND = 268 };
typedef union YYSTYPE {
    double num;
    const char *str;
    struct INPparseNode *pnode;
} YYSTYPE;
#endif
