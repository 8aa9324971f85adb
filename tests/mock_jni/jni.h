/* Minimal stand-in for jni.h, only so that tests can syntax-check integration/jni/ubu_sw_jni.c against include/*.h (no JDK in this image). */
#include <stdint.h>
#include <stdio.h>
typedef int32_t jint; typedef int64_t jlong; typedef void* jobject; typedef jobject jclass; typedef jobject jstring; typedef unsigned char jboolean;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
  jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  jstring (*NewStringUTF)(JNIEnv*, const char*);
};
#define JNIEXPORT
#define JNICALL
