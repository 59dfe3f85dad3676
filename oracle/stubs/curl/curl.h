/* TEST INFRASTRUCTURE — stand-in for <curl/curl.h> (libcurl headers are not
 * installed; there is no network either).  Covers the handful of names the
 * reference's FTP download block uses (plutogpssim.c:2230-2231, 2447-2468).
 * curl_easy_perform always reports failure, so "-f" exits like a failed fetch. */
#ifndef ORACLE_STUB_CURL_H
#define ORACLE_STUB_CURL_H

typedef void CURL;
typedef enum { CURLE_OK = 0, CURLE_GOT_NOTHING = 52 } CURLcode;
typedef enum {
    CURLOPT_URL = 10002,
    CURLOPT_WRITEFUNCTION = 20011,
    CURLOPT_WRITEDATA = 10001,
    CURLOPT_USE_SSL = 119,
    CURLOPT_VERBOSE = 41,
    CURLOPT_USERPWD = 10005
} CURLoption;
enum { CURLUSESSL_NONE = 0 };
#define CURL_GLOBAL_DEFAULT 3

CURLcode curl_global_init(long flags);
void curl_global_cleanup(void);
CURL *curl_easy_init(void);
CURLcode curl_easy_setopt(CURL *h, CURLoption opt, ...);
CURLcode curl_easy_perform(CURL *h);
void curl_easy_cleanup(CURL *h);

#endif
